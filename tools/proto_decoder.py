"""Numerics prototype (numpy float32) of the GPU decoder formulation: VN in likelihood-ratio domain,
CN in (hard bit, small probability) domain.  Compares with the fp64 oracle.  Dev tool only."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kml_oracle as ko

f32 = np.float32

def build(code):
    ex = code.export(with_enc=False)
    rp, ci = ex["row_ptr"], ex["col_idx"]
    M, N, E = code.M, code.N, code.E
    edge_col = ci.copy()
    edge_row = np.repeat(np.arange(M), np.diff(rp))
    # per variable list of edges
    order = np.argsort(edge_col, kind="stable")
    cp = np.zeros(N + 1, np.int64); np.add.at(cp, edge_col + 1, 1); cp = np.cumsum(cp)
    return rp, edge_col, edge_row, order, cp

def decode_lr(code, p0, iters, max_iter):
    """p0: [B, N_tx] float64 from oracle. returns ret[B], cc_hat[B,N]"""
    rp, edge_col, edge_row, order, cp = build(code)
    B = p0.shape[0]; M, N, E = code.M, code.N, code.E
    p0f = p0.astype(f32)
    # channel LR computed from (p0, 1-p0) in fp64 then rounded: the GPU demapper produces z0/z1 directly
    m = np.ones((B, N), f32)
    m[:, code.two_z:] = (p0 / (1.0 - p0)).astype(f32)
    x = np.ones((B, E), f32)            # c2v LR
    done = np.zeros(B, bool); ret = np.full(B, iters, np.int32)
    out_bits = np.zeros((B, N), np.int8)
    deg_v = np.diff(cp); maxdv = deg_v.max()
    deg_c = np.diff(rp); maxdc = deg_c.max()
    CL = f32(1e-12)
    for t in range(iters):
        # VN: total product via log? emulate sequential multiply in fp32 per variable: do prefix/suffix
        # gather c2v per variable padded with 1
        idx = np.full((N, maxdv), E, np.int64)
        for v in range(N):
            idx[v, :deg_v[v]] = order[cp[v]:cp[v+1]]
        xp = np.concatenate([x, np.ones((B, 1), f32)], axis=1)
        g = xp[:, idx]                                   # B,N,maxdv
        with np.errstate(over="ignore", under="ignore", invalid="ignore"):
            # prefix products including channel
            pre = np.empty((B, N, maxdv + 1), f32); pre[:, :, 0] = m
            for k in range(maxdv):
                pre[:, :, k + 1] = np.clip(pre[:, :, k] * g[:, :, k], f32(1e-36), f32(1e36))
            suf = np.empty((B, N, maxdv + 1), f32); suf[:, :, maxdv] = 1
            for k in range(maxdv - 1, -1, -1):
                suf[:, :, k] = np.clip(suf[:, :, k + 1] * g[:, :, k], f32(1e-36), f32(1e36))
            post = pre[:, :, maxdv]
            bits = (~(post > 1)).astype(np.int8)
            ext = np.clip(pre[:, :, :maxdv] * suf[:, :, 1:], f32(1e-36), f32(1e36))  # B,N,maxdv
            hard = (ext < 1)
            s = (np.minimum(ext, f32(1)) / (f32(1) + ext)).astype(f32)
        # syndrome
        synd = np.zeros((B, M), np.int8)
        np.bitwise_xor.at(synd, (slice(None), edge_row), bits[:, edge_col])
        ok = ~synd.any(axis=1)
        newly = ok & ~done
        first = ~done
        out_bits[first] = bits[first]
        ret[newly] = t + 1 if t < max_iter else t
        done |= ok
        if done.all(): break
        # scatter v2c to edges
        hs = np.zeros((B, E + 1), bool); ss = np.zeros((B, E + 1), f32)
        hs[:, idx] = hard; ss[:, idx] = s
        hs = hs[:, :E]; ss = ss[:, :E]
        # CN with padding to maxdc: neutral element s=0,h=0
        cidx = np.full((M, maxdc), E, np.int64)
        for r in range(M):
            cidx[r, :deg_c[r]] = np.arange(rp[r], rp[r+1])
        hp = np.concatenate([hs, np.zeros((B, 1), bool)], 1)[:, cidx]
        sp = np.concatenate([ss, np.zeros((B, 1), f32)], 1)[:, cidx]
        def op(a, b): return (a + b - f32(2) * a * b).astype(f32)
        pre = np.zeros((B, M, maxdc + 1), f32)
        for k in range(maxdc): pre[:, :, k + 1] = op(pre[:, :, k], sp[:, :, k])
        suf = np.zeros((B, M, maxdc + 1), f32)
        for k in range(maxdc - 1, -1, -1): suf[:, :, k] = op(suf[:, :, k + 1], sp[:, :, k])
        so = op(pre[:, :, :maxdc], suf[:, :, 1:])
        so = np.maximum(so, CL)
        H = np.logical_xor.reduce(hp, axis=2)
        ho = H[:, :, None] ^ hp
        with np.errstate(over="ignore"):
            lr = np.where(ho, so / (f32(1) - so), (f32(1) - so) / so).astype(f32)
        xn = np.ones((B, E + 1), f32)
        xn[:, cidx] = lr
        x = np.where(done[:, None], x, xn[:, :E])
    return ret, out_bits

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "peg2304_4psk_6db"
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".npz"))
    p = json.loads(str(z["params"]))
    link = ko.Link(p["matrix"], p["modem"], is_5g=bool(p["g5"]), active=bool(p["active"]), known_h=bool(p["known_h"]),
                   metric_type=bool(p["metric_type"]), metric_iter=p["metric_iter"], max_iter=p["max_iter"])
    g = ko.Lcg(17)
    F = int(sys.argv[2]) if len(sys.argv) > 2 else p["frames"]
    rs = [link.frame(g, p["snr"]) for _ in range(F)]
    p0 = np.stack([r.p0 for r in rs])
    ret, bits = decode_lr(link.code, p0, p["max_iter"], p["max_iter"])
    oret = np.array([r.ret for r in rs]); obits = np.stack([r.cc_hat for r in rs])
    conv = oret < p["max_iter"]   # ret==max_iter may also be converged at last iter; approximate
    same_ret = (ret == oret)
    same_bits = (bits == obits).all(axis=1)
    print(name, "frames", F, "ret identical", same_ret.sum(), "/", F,
          "| converged", conv.sum(), "bit-identical among converged", (same_bits & conv).sum(),
          "| non-converged identical", (same_bits & ~conv).sum(), "/", (~conv).sum())
    bad = np.where(~same_ret)[0]
    print("ret mismatches:", [(int(i), int(ret[i]), int(oret[i])) for i in bad[:10]])
