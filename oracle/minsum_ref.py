"""TEST INFRASTRUCTURE ONLY — numpy (float32) restatement of the normalised min-sum flooding decoder offered as
`algorithm = 1` (throughput mode).  NOT reference-pinned: the reference has no min-sum; this file
exists so the CUDA min-sum kernel has an independent statement of the same algorithm to be checked against, and the
algorithm itself is gated by BER/FER against the sum-product oracle (tests/test_gpu_minsum.py).

Schedule, stopping rule and return value are the reference decoder's (binaryldpccodec.cc:175-277); only the node
updates differ:  VN  total = L_ch + sum c2v,  v2c_e = total - c2v_e,  bit = (total > 0) ? 0 : 1
                 CN  c2v_e = max(alpha * min_{e' != e} |v2c_e'| - beta, 0) * prod sign,  clipped to +-ln((1-1e-12)/1e-12)
                     (normalised: beta = 0; offset: alpha = 1, beta > 0)."""
from __future__ import annotations

import numpy as np

f32 = np.float32
LLR_CLIP = f32(27.631021)


def decode(row_ptr, col_idx, n_graph, punct, llr, iters, max_iter, alpha=0.8, beta=0.0):
    """llr [B, n_tx] float32 (ln P0/P1).  Returns ret[B], cc_hat[B, n_graph]."""
    rp, ci = np.asarray(row_ptr), np.asarray(col_idx)
    M, E = len(rp) - 1, len(ci)
    B = llr.shape[0]
    edge_row = np.repeat(np.arange(M), np.diff(rp))
    ch = np.zeros((B, n_graph), f32)
    ch[:, punct:] = np.clip(llr.astype(f32), -LLR_CLIP, LLR_CLIP)
    c2v = np.zeros((B, E), f32)
    ret = np.full(B, iters + (1 if iters < max_iter else 0), np.int32)
    done = np.zeros(B, bool)
    out = np.zeros((B, n_graph), np.int8)
    alpha, beta = f32(alpha), f32(beta)
    for t in range(iters):
        total = ch.copy()
        np.add.at(total, (slice(None), ci), c2v)
        bits = (~(total > 0)).astype(np.int8)
        synd = np.zeros((B, M), np.int8)
        np.bitwise_xor.at(synd, (slice(None), edge_row), bits[:, ci])
        ok = ~synd.any(axis=1)
        out[~done] = bits[~done]
        ret[ok & ~done] = t + (1 if t < max_iter else 0)
        done |= ok
        if done.all():
            break
        v2c = (total[:, ci] - c2v).astype(f32)
        mag = np.abs(v2c)
        neg = v2c < 0  # sign bit; -0.0 cannot occur because total - c2v of finite values
        new = np.zeros_like(c2v)
        for r in range(M):
            a, b = rp[r], rp[r + 1]
            m = mag[:, a:b]
            srt = np.sort(m, axis=1)
            min1, min2 = srt[:, :1], srt[:, 1:2] if b - a > 1 else srt[:, :1]
            par = np.logical_xor.reduce(neg[:, a:b], axis=1, keepdims=True)
            o = np.where(m == min1, min2, min1).astype(np.float64) * np.float64(alpha) - np.float64(beta)  # (one rounding, like fmaf)
            o = np.minimum(np.maximum(o.astype(f32), f32(0)), LLR_CLIP)
            sgn = par ^ neg[:, a:b]
            new[:, a:b] = np.where(sgn, -o, o)
        c2v = np.where(done[:, None], c2v, new).astype(f32)
    return ret, out


def qc_structure(row_ptr, col_idx, n_rows, n_cols):
    """Largest Z for which H is a grid of Z x Z blocks that are zero or one cyclically shifted identity.
    Returns (Z, layers) with layers[l] = [(block column, shift), ...] sorted by column, or (0, None)."""
    rp, ci = np.asarray(row_ptr), np.asarray(col_idx)
    g = int(np.gcd(n_rows, n_cols))
    for Z in range(g, 7, -1):
        if g % Z:
            continue
        layers, ok = [], True
        for l in range(n_rows // Z):
            first = None
            for z in range(Z):
                r = l * Z + z
                mine = sorted((int(c) // Z, (int(c) % Z - z) % Z) for c in ci[rp[r]:rp[r + 1]])
                if len({b for b, _ in mine}) != len(mine):
                    ok = False
                if first is None:
                    first = mine
                elif mine != first:
                    ok = False
                if not ok:
                    break
            if not ok:
                break
            layers.append(first)
        if ok:
            return Z, layers
    return 0, None


def decode_layered(layers, Z, n_graph, punct, llr, iters, max_iter, alpha=0.8, beta=0.0):
    """Layered (row-serial) min-sum for a quasi-cyclic code, `algorithm = 3` (kmldpc_b200/csrc/bp_layered.cu): block rows in
    order, every check seeing the posteriors as the layers before it left them.  Old messages are kept as (min1, min2)
    rounded to float16, so what a visit subtracts is exactly what the last one added.  Stops after the first iteration in
    which every check saw satisfied parity and no decision moved.  Returns ret[B], cc_hat[B, n_graph]."""
    B = llr.shape[0]
    L = np.zeros((B, n_graph), f32)
    L[:, punct:] = np.clip(llr.astype(f32), -LLR_CLIP, LLR_CLIP)
    zz = np.arange(Z)
    cols = [np.stack([b * Z + (zz + s) % Z for b, s in lay]) for lay in layers]  # [d, Z] per layer
    old = [np.zeros((B,) + c.shape, f32) for c in cols]
    ret = np.full(B, iters + (1 if iters < max_iter else 0), np.int32)
    done = np.zeros(B, bool)
    out = np.zeros((B, n_graph), np.int8)
    a64, b64 = np.float64(f32(alpha)), np.float64(f32(beta))

    def scale(m):
        s = (m.astype(np.float64) * a64 - b64).astype(f32)  # one rounding, like fmaf
        s = np.minimum(np.maximum(s, f32(0)), LLR_CLIP)
        return s.astype(np.float16).astype(f32)

    for t in range(iters):
        fail = np.zeros(B, bool)
        for l, c in enumerate(cols):
            lv = L[:, c]                                  # [B, d, Z]
            dec = ~(lv > 0)
            v = (lv - old[l]).astype(f32)
            a, neg = np.abs(v), v < 0
            idx = np.argmin(a, axis=1)                    # first minimum, like the kernel's strict compare
            m1 = np.take_along_axis(a, idx[:, None, :], axis=1)[:, 0, :]
            a2 = a.copy()
            np.put_along_axis(a2, idx[:, None, :], np.inf, axis=1)
            m2 = a2.min(axis=1) if a.shape[1] > 1 else np.full_like(m1, 3.0e38)
            q1, q2 = scale(m1), scale(np.minimum(m2, f32(3.0e38)))
            is_min = np.arange(a.shape[1])[None, :, None] == idx[:, None, :]
            mag = np.where(is_min, q2[:, None, :], q1[:, None, :])
            sg = np.logical_xor.reduce(neg, axis=1, keepdims=True) ^ neg
            new = np.where(sg, -mag, mag).astype(f32)
            nl = (v + new).astype(f32)
            fail |= ((~(nl > 0)) != dec).any(axis=(1, 2)) | np.logical_xor.reduce(dec, axis=1).any(axis=1)
            upd = ~done
            L[np.ix_(upd, c.ravel())] = nl[upd].reshape(upd.sum(), -1)
            old[l][upd] = new[upd]
        ok = ~fail & ~done
        ret[ok] = t + 1 + (1 if t + 1 < max_iter else 0)
        done |= ok
        if done.all():
            break
    out[:] = ~(L > 0)
    return ret, out
