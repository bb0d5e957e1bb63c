#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_harness, built from
/root/reference by `make -C oracle ref`).  Run in the build container only; the fixtures are committed.

Each case writes
  <name>.npz : run parameters, `full` tensors for the first FULL frames (u, c, h, y, clusters, p0, cc_hat, uu_hat …)
               and per-frame summaries for ALL frames (h, hhat, metric[4], kstar, ret, nerr, checksums).
  code_<matrix>.npz : the reference object's Tanner graph (traversal-order CSR), code parameters and a SHA-256 of
               the reduced encoder matrix enc_h_ (plus a few full rows).
The global LCG is seeded with SetSeed(-1) (state 17) and frames are drawn in the reference's own order, so the C
restatement regenerates the same frames and must reproduce every value here.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
CFG = os.path.join(ROOT, "config")

PEG = "PEG2304regular0.5.txt"
PEG8 = "PEG8064regular0.5.txt"
G5 = "5GLDPCBG2a3_R12_K960.txt"

# name, matrix, modem, snr, frames, full, extra
CASES = [
    ("peg2304_qpsk_10db", PEG, "2bits_QPSK.txt", 10.0, 200, 4, {}),
    ("peg2304_qpsk_m5db", PEG, "2bits_QPSK.txt", -5.0, 40, 2, {}),
    ("peg2304_4psk_6db", PEG, "2bits_4PSK.txt", 6.0, 200, 4, {}),
    ("peg2304_16qam_gray_12db", PEG, "4bit_16QAM_Gray.txt", 12.0, 200, 4, {}),
    ("peg2304_16qam_phi1_15db", PEG, "4bit_16QAM_phi1.txt", 15.0, 40, 2, {}),
    ("peg2304_16qam_phi2_known_15db", PEG, "4bit_16QAM_phi2.txt", 15.0, 100, 2, {"known_h": 1}),
    ("5g_16qam_gray_10db", G5, "4bit_16QAM_Gray.txt", 10.0, 100, 4, {"g5": 1, "metric_iter": 5}),
    ("peg8064_64qam_20db", PEG8, "6bits_64QAM_Gray.txt", 20.0, 12, 2, {}),
    ("peg2304_4psk_soft_6db", PEG, "2bits_4PSK.txt", 6.0, 40, 2, {"metric_type": 1, "metric_iter": 5}),
    # soft metric where most correct candidates leave at iteration 0: their metric is the STALE syndrom_soft_ of the
    # previous Decoder call (binaryldpccodec.cc:231-232,274), carried from candidate to candidate and frame to frame
    ("peg2304_4psk_soft_18db", PEG, "2bits_4PSK.txt", 18.0, 120, 2, {"metric_type": 1, "metric_iter": 5}),
    ("peg2304_4psk_inactive_6db", PEG, "2bits_4PSK.txt", 6.0, 40, 2, {"active": 0}),
    ("peg2304_16qam_gray_i5_15db", PEG, "4bit_16QAM_Gray.txt", 15.0, 100, 2, {"max_iter": 5}),
]


# histogram mode ([histogram] enable = true) through the reference's own KmCodec::GetHistogramData: same LCG frames as
# the case of the same name without the prefix; name, matrix, modem, snr, frames, extra
HIST_CASES = [
    ("hist_peg2304_4psk_6db", PEG, "2bits_4PSK.txt", 6.0, 100, {}),
    ("hist_5g_16qam_gray_10db", G5, "4bit_16QAM_Gray.txt", 10.0, 60, {"g5": 1, "metric_iter": 5}),
]


def rd(d, name, dtype, shape=None):
    p = os.path.join(d, name)
    if not os.path.exists(p):
        return None
    a = np.fromfile(p, dtype=dtype)
    return a.reshape(shape) if shape is not None else a


def pack(bits):
    return np.packbits(bits.astype(np.uint8), axis=-1)


def run_case(name, matrix, modem, snr, frames, full, extra):
    with tempfile.TemporaryDirectory() as d:
        args = dict(cfgdir=CFG, matrix=matrix, modem=modem, snr=snr, frames=frames, out=d, g5=0, active=1,
                    known_h=0, metric_type=0, metric_iter=5, max_iter=50)
        args.update(extra)
        cmd = [HARNESS, "dump"] + [f"{k}={v}" for k, v in args.items()]
        out = subprocess.check_output(cmd, text=True)
        summ = json.loads(out.strip().splitlines()[-1])
        assert summ["kmcodec_mismatch"] == 0, "harness mirror disagrees with KmCodec::Decoder"
        n_tx, n_graph, k, m, q, n_sym, two_z, _ = rd(d, "run_meta.i32", np.int32)
        F = frames
        u = rd(d, "u.i8", np.int8, (F, k)); c = rd(d, "c.i8", np.int8, (F, n_tx))
        h = rd(d, "h.f64", np.float64, (F, 2)); y = rd(d, "y.f64", np.float64, (F, n_sym, 2))
        cl = rd(d, "clusters.f64", np.float64, (F, q, 2)); hhat = rd(d, "hhat.f64", np.float64, (F, 2))
        metric = rd(d, "metric.f64", np.float64, (F, 4)); kstar = rd(d, "kstar.i32", np.int32)
        p0 = rd(d, "p0.f64", np.float64, (F, n_tx)); cch = rd(d, "cc_hat.i8", np.int8, (F, n_graph))
        uh = rd(d, "uu_hat.i8", np.int8, (F, k)); ret = rd(d, "ret.i32", np.int32); nerr = rd(d, "nerr.i32", np.int32)
        cons = rd(d, "constellation.f64", np.float64, (q, 2))
        fixture = dict(
            params=json.dumps(dict(name=name, matrix=matrix, modem=modem, snr=snr, frames=F, full=full, **{
                kk: args[kk] for kk in ("g5", "active", "known_h", "metric_type", "metric_iter", "max_iter")},
                n_tx=int(n_tx), n_graph=int(n_graph), k=int(k), m=int(m), q=int(q), n_sym=int(n_sym),
                two_z=int(two_z), ber=summ["ber"], fer=summ["fer"])),
            constellation=cons,
            # summaries for every frame
            h=h, hhat=hhat, metric=metric, kstar=kstar, ret=ret, nerr=nerr,
            y_sum=y.sum(axis=1), p0_sum=p0.sum(axis=1), c_weight=c.sum(axis=1).astype(np.int32),
            cc_hat_weight=cch.sum(axis=1).astype(np.int32), clusters0=cl[:, 0, :],
            uu_hat_packed=pack(uh),
            # full tensors for the first `full` frames
            f_u=pack(u[:full]), f_c=pack(c[:full]), f_y=y[:full], f_clusters=cl[:full], f_p0=p0[:full],
            f_cc_hat=pack(cch[:full]), f_uu_hat=pack(uh[:full]),
        )
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fixture)
        # code fixture (once per matrix/active flavour)
        tag = matrix.replace(".txt", "") + ("" if args["active"] else "_inactive")
        cpath = os.path.join(HERE, "code_" + tag + ".npz")
        if not os.path.exists(cpath):
            meta = rd(d, "code_meta.i32", np.int32)
            enc = rd(d, "enc_h.i8", np.int8)
            code = dict(meta=meta, row_ptr=rd(d, "row_ptr.i32", np.int32), col_idx=rd(d, "col_idx.i32", np.int32),
                        col_ptr=rd(d, "col_ptr.i32", np.int32), row_idx=rd(d, "row_idx.i32", np.int32))
            if enc is not None:
                enc = enc.reshape(int(meta[0]), int(meta[1]))
                code["enc_sha256"] = np.frombuffer(hashlib.sha256(enc.tobytes()).digest(), np.uint8)
                code["enc_rows_0_1_last"] = np.packbits(enc[[0, 1, -1]].astype(np.uint8), axis=-1)
                code["enc_row_weight"] = enc.sum(axis=1).astype(np.int32)
            np.savez_compressed(cpath, **code)
        print(f"{name}: FER {summ['fer']:.3f} BER {summ['ber']:.4f} avg_ret {summ['avg_ret']:.1f} "
              f"-> {os.path.getsize(os.path.join(HERE, name + '.npz')) / 1024:.0f} KiB")


def run_hist_case(name, matrix, modem, snr, frames, extra):
    with tempfile.TemporaryDirectory() as d:
        args = dict(cfgdir=CFG, matrix=matrix, modem=modem, snr=snr, frames=frames, out=d, g5=0, active=1,
                    known_h=0, metric_type=0, metric_iter=5, max_iter=50, hist=1)
        args.update(extra)
        out = subprocess.check_output([HARNESS, "dump"] + [f"{k}={v}" for k, v in args.items()], text=True)
        summ = json.loads(out.strip().splitlines()[-1])
        n_tx, n_graph, k, m, q, n_sym, two_z, _ = rd(d, "run_meta.i32", np.int32)
        F = frames
        fixture = dict(
            params=json.dumps(dict(name=name, matrix=matrix, modem=modem, snr=snr, frames=F, histogram=1, **{
                kk: args[kk] for kk in ("g5", "active", "known_h", "metric_type", "metric_iter", "max_iter")},
                n_tx=int(n_tx), k=int(k), tot_blk=summ["tot_blk"], err_blk=summ["err_blk"], ber=summ["ber"], fer=summ["fer"])),
            h=rd(d, "h.f64", np.float64, (F, 2)), hhat=rd(d, "hhat.f64", np.float64, (F, 2)),
            metric=rd(d, "metric.f64", np.float64, (F, 4)), hist_line=rd(d, "hist_line.f64", np.float64, (F, 4)),
            kstar=rd(d, "kstar.i32", np.int32), nerr=rd(d, "nerr.i32", np.int32),
            uu_hat_packed=pack(rd(d, "uu_hat.i8", np.int8, (F, k))))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fixture)
        print(f"{name}: histogram mode, FER {summ['fer']:.3f} BER {summ['ber']:.4f}")


if __name__ == "__main__":
    if not os.path.exists(HARNESS):
        sys.exit("build oracle/_ref first: make -C oracle ref   (needs /root/reference)")
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        run_case(*case)
    for case in HIST_CASES:
        if only and case[0] not in only:
            continue
        run_hist_case(*case)
