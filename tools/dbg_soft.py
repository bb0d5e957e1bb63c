"""Debug: soft-metric chain of the GPU receiver against the oracle, frame by frame (one chain of F frames)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util
from oracle import kml_oracle as ko
name, F = sys.argv[1], int(sys.argv[2])
snr = util.CASES[name][2]
olink = util.oracle_link(name)
g = ko.Lcg(17)
rs = [olink.frame(g, snr, full=True) for _ in range(F)]
link = util.gpu_link(name, max_batch=4096)
y = np.stack([r.y for r in rs])
uu_p, hhat, kstar, ret, met = link.receive_f64(y, 10 ** (-0.1 * snr), with_metric=True)
ref_m = np.stack([r.metric for r in rs]); ref_k = np.array([r.kstar for r in rs]); ref_ret = np.array([r.ret for r in rs])
bad = np.where(kstar != ref_k)[0]
rel = np.abs(met - ref_m) / np.maximum(ref_m, 1e-9)
print("frames", F, "kstar mismatches", len(bad), "ret mismatches", int((ret != ref_ret).sum()))
print("metric rel err: median %.2e p99 %.2e max %.2e, inf count gpu %d" % (np.median(rel), np.quantile(rel, 0.99), rel.max(), int(np.isinf(met).sum())))
worst = np.argsort(-rel.max(axis=1))[:8]
for f in sorted(set(list(bad[:10]) + list(worst))):
    print(f, "gpu", met[f], kstar[f], ret[f], "| ref", ref_m[f], ref_k[f], ref_ret[f], "| prev ref", ref_m[f - 1] if f else None, ref_k[f - 1] if f else None, ref_ret[f - 1] if f else None)
