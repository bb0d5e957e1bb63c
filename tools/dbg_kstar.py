"""Debug: metrics / k* of the GPU resolver against the oracle for one test case (which frames differ, and by how much)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util
name, frames = sys.argv[1], int(sys.argv[2])
olink, rs = util.oracle_frames(name, frames)
link = util.gpu_link(name, max_batch=64)
var = 10 ** (-0.1 * util.CASES[name][2])
y = np.stack([r.y for r in rs])
ref_h = np.array([r.hhat for r in rs])
h64, _ = link.kmeans_f64(y)
h32, _ = link.kmeans(y.astype(np.complex64))
print("hhat rel f64", (np.abs(h64 - ref_h) / np.abs(ref_h)).max(), "f32", (np.abs(h32 - ref_h) / np.abs(ref_h)).max())
for tag, hh in (("ref hhat", ref_h), ("gpu hhat32", h32)):
    met, ks = link.resolve(y, hh, var)
    ref_m = np.stack([r.metric for r in rs]); ref_k = np.array([r.kstar for r in rs])
    bad = np.where(ks != ref_k)[0]
    print(tag, "kstar mismatches", bad, "metric rows differing", np.where((met != ref_m).any(axis=1))[0])
    for f in bad:
        print("  frame", f, "gpu", met[f], ks[f], "ref", ref_m[f], ref_k[f], "ret", rs[f].ret)
uu_p, hhat, kstar, ret = link.receive(y, var)
print("receive kstar", kstar, "ref", np.array([r.kstar for r in rs]))
