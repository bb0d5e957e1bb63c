import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kmldpc_b200 as kb
from tests import util
name = sys.argv[1] if len(sys.argv) > 1 else "peg8064_64qam_20db"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
olink, rs = util.oracle_frames(name, frames)
link = util.gpu_link(name, max_batch=64)
var = 10 ** (-0.1 * util.CASES[name][2])
y = np.stack([r.y for r in rs])
uu_p, hhat, kstar, ret = link.receive(y, var)
uu = kb.unpack_bits(uu_p, olink.code.K)
ref_uu = np.stack([r.uu_hat for r in rs]); ref_ret = np.array([r.ret for r in rs]); ref_k = np.array([r.kstar for r in rs])
syn = np.array([olink.code.parity_check(r.cc_hat) for r in rs])
print("ret ", ret); print("rret", ref_ret); print("k", kstar, ref_k); print("syn", syn)
print("diff bits per frame", (uu != ref_uu).sum(axis=1))
ref_h = np.array([r.hhat for r in rs]); print("hhat rel", np.abs(hhat-ref_h)/np.abs(ref_h))
# stage by stage
rot = np.exp(1j * (3.14159265358979 / 2) * np.arange(4))
llr = link.demap(y, hhat * rot[kstar], var)
ref_llr = np.stack([util.llr_of_p0(r.p0) for r in rs])
print("llr maxerr per frame", np.abs(llr - ref_llr).max(axis=1))
cc, uu2, ret2 = link.decode(llr)
print("decode(llr from gpu demap) ret", ret2, "diff", (uu2 != ref_uu).sum(axis=1))
cc, uu3, ret3 = link.decode(ref_llr.astype(np.float32))
print("decode(ref llr) ret", ret3, "diff", (uu3 != ref_uu).sum(axis=1))
