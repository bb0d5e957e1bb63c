#!/usr/bin/env python
"""bench.py — decoded Mbit/s of the kmldpc link receiver on B200 (and the reference CPU arm).

Workload (config.workload "C1-fixedI50"): BASELINE.json configs[0] — PEG2304 regular R1/2 + 2bits_QPSK.txt, blind
k-means (20 passes) → 4-rotation resolve → soft demap → flooding sum-product BP, max_iter = 50 — at the FIXED iteration
count the metric is quoted on: SNR = -5 dB, where the syndrome never clears, so every frame executes exactly 50
iterations in BOTH arms with the reference's semantics unchanged.  The reported
`iters_per_frame` proves it.

One step = one batch of B frames through the receiver path (k-means → resolve → demap → decode → error count).
  value : inputs (received symbols y, float32) already resident in HBM, kml_receive_dev + kml_count_errors_dev on one
          stream, CUDA events on that stream.  Batches rotate over a pool larger than L2.
  e2e   : the same batches through the host-buffer C-ABI call kml_receive (pinned host y in → packed decisions out),
          H2D + D2H inside the timed region.
  roofline : the BP decoder kernel alone (kml_decode_dev on LLRs resident in HBM), CUDA events, against the shared
          memory bandwidth bound of SURVEY §8(d) (16 B per edge-iteration); HBM figures for its LLR I/O beside it.
  cpu_baseline / --impl reference : the UNMODIFIED reference classes (oracle/_ref/ref_harness, built from
          /root/reference in the build container) on all host cores, receiver stages only.  A missing harness is an
          ERROR (the C port is used only when --ref-kind port asks for it).  cpu_baseline.stock_binary = the reference's
          own multi-threaded binary (oracle/_ref/kmldpc_stock: its unchanged main + Simulator + thread pool) on the same
          workload, wall time from its own "Total time cost" line minus a 1-frame setup run.
Extra legs (rank 0, after the headline):
  configs : BASELINE configs 2-4 (5G BG2 + 16QAM, PEG8064 + 64QAM, PEG2304 + 16QAM phi1) — receiver value, decoder
          roofline and CPU baseline each, at a fixed 50 iterations (N = 1 only).
  c5      : BASELINE config 5 — the full 31-point SNR sweep, 10^8 frames, through the product's own multi-GPU driver
          kml_sweep_run on all N GPUs of the job (one host thread per GPU, ONE ncclAllReduce of the counters per point):
          fixed total work = STRONG scaling; counters_checksum must be identical for every N.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MATRIX, MODEM = "PEG2304regular0.5.txt", "2bits_QPSK.txt"
SNR_DB, MAX_ITER, KMEANS_ITER = -5.0, 50, 20
K_INFO, N_CODE, N_EDGES, N_SYM = 1152, 2304, 6912, 1152
METRIC = "LDPC decoded Mbit/s (PEG2304 R1/2 + QPSK, blind k-means detect + BP decode, fixed 50 iterations)"
CONFIG = {"workload": "C1-fixedI50", "code": "PEG2304 regular R1/2", "modem": "2bits_QPSK", "snr_db": SNR_DB,
          "max_iter": MAX_ITER, "kmeans_iter": KMEANS_ITER, "detector": "blind k-means + 4-rotation resolve",
          "decoder": "flooding sum-product (the reference's algorithm)",
          "path": "receiver: y -> k-means -> resolve -> demap -> BP -> error count"}


def make_config(args):
    """The same dict in both arms (the workload is the same; batch / pool describe how the GPU arm steps through it)."""
    return dict(CONFIG, batch_frames_per_gpu=args.batch, input_pool_batches=args.pool,
                l2_policy=f"inputs rotate over {args.pool} batches = {args.pool * args.batch * N_SYM * 8 / 2**20:.0f} MiB > 126 MiB L2")



# ------------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [r[1:] for r in self.rows if len(r) >= 10 and r[2].replace(".", "").isdigit()]
        if self.t_begin is not None and self.t_end is not None:  # samples taken while the timed legs were running
            inside = [r[1:] for r in self.rows if len(r) >= 10 and r[2].replace(".", "").isdigit()
                      and self.t_begin <= r[0] <= self.t_end]
            if len(inside) >= 3:
                rows = inside
        sm = [float(r[1]) for r in rows]
        mx = [float(r[2]) for r in rows]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None,
                "window": "value + e2e + roofline legs (GPU under load)"}


# ------------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference_sample(frames_per_proc: int, procs: int, seed0: int = 1000, kind: str = "reference", *, matrix=MATRIX,
                         modem=MODEM, snr_db=SNR_DB, k_info=K_INFO, is_5g=False):
    """Runs the reference's own receiver code on `procs` host processes.  Returns dict with decoded Mbit/s of the
    receiver stages (k-means + resolve + demap + decode), whole-frame Mbit/s, kind, cores, iterations per frame."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    cfg = os.path.join(ROOT, "config")
    if kind == "reference":
        if not (os.path.exists(harness) and os.access(harness, os.X_OK)):
            raise SystemExit("bench.py: oracle/_ref/ref_harness is missing — build it where /root/reference exists "
                             "(make -C oracle ref; it travels with the repo), or pass --ref-kind port to time the C port")
        cmds = [[harness, "time", f"cfgdir={cfg}", f"matrix={matrix}", f"modem={modem}", f"snr={snr_db}", f"g5={int(is_5g)}",
                 f"frames={frames_per_proc}", f"max_iter={MAX_ITER}", f"seed={seed0 + 7919 * i}"] for i in range(procs)]
        t0 = time.perf_counter()
        ps = [subprocess.Popen(c, stdout=subprocess.PIPE, text=True) for c in cmds]
        outs = [p.communicate()[0] for p in ps]
        wall = time.perf_counter() - t0
        rs = [json.loads(o.strip().splitlines()[-1]) for o in outs]
        rx = [r["t_kmeans"] + r["t_resolve"] + r["t_demap"] + r["t_decode"] for r in rs]
        full = [x + r["t_src_enc"] + r["t_chan"] for x, r in zip(rx, rs)]
        frames = frames_per_proc * procs
        return {"kind": "reference", "cores": procs, "frames": frames, "wall_s": wall,
                "rx_mbps": frames * k_info / max(rx) / 1e6, "frame_mbps": frames * k_info / max(full) / 1e6,
                "rx_s": max(rx), "iters_per_frame": statistics.mean(r["avg_ret"] for r in rs),
                "ms_per_frame_per_core": 1e3 * statistics.mean(rx) / frames_per_proc}
    # C port of the same algorithm (oracle/kml_oracle.c), pthreads; whole frame (its stages are not timed apart)
    from oracle import kml_oracle as ko
    link = ko.Link(matrix, modem, max_iter=MAX_ITER, is_5g=is_5g)
    frames = frames_per_proc * procs
    t0 = time.perf_counter()
    cnt, iters = link.run(snr_db, frames, threads=procs, seed0=seed0)
    wall = time.perf_counter() - t0
    return {"kind": "port", "cores": procs, "frames": frames, "wall_s": wall, "rx_mbps": frames * k_info / wall / 1e6,
            "frame_mbps": frames * k_info / wall / 1e6, "rx_s": wall, "iters_per_frame": iters / frames,
            "ms_per_frame_per_core": 1e3 * wall * procs / frames}


def run_stock_binary(frames: int, threads: int):
    """The reference's own binary on the headline workload: its unchanged main(), Simulator, thread pools and logger,
    reading a config.toml (kmldpc.cpp:29) in a scratch directory.  Wall time from its own 'Total time cost' line
    (kmldpc.cpp:44-53) minus that of a 1-frame run (setup: H file + Gaussian elimination).  Its RNG is seeded from
    time() and shared unlocked between threads, so frames differ from run to run; at -5 dB every frame runs 50 iterations."""
    import re
    import shutil
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "kmldpc_stock")
    if not (os.path.exists(exe) and os.access(exe, os.X_OK)):
        return {"unavailable": "oracle/_ref/kmldpc_stock not built (make -C oracle ref where /root/reference exists)"}

    def once(n):
        with tempfile.TemporaryDirectory() as d:
            for f in (MATRIX, MODEM):
                shutil.copy(os.path.join(ROOT, "config", f), d)
            os.makedirs(os.path.join(d, "logs"))
            per = max(1, -(-n // threads))
            open(os.path.join(d, "config.toml"), "w").write(
                f"[range]\nminimum_snr = {SNR_DB}\nmaximum_snr = {SNR_DB}\nstep_snr = 1.0\nmaximum_error_number = 100000000\n"
                f"maximum_block_number = {n}\nthread_block_number = {per}\n[decoder]\ntrue_h_arg = false\n[xcodec]\n5gldpc = false\n"
                f"metric_type = false\nmetric_iter = 5\n[histogram]\nenable = false\n[ldpc]\nmax_iter = {MAX_ITER}\nactive = true\n"
                f"matrix_file = \"{MATRIX}\"\n[modem]\nmodem_file = \"{MODEM}\"\n")
            t0 = time.perf_counter()
            out = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=600).stdout
            wall = time.perf_counter() - t0
            m = re.search(r"Total time cost: (\d+)min:(\d+)sec:(\d+)ms", out)
            own = (int(m.group(1)) * 60000 + int(m.group(3))) / 1e3 if m else wall  # (its ms field already holds the seconds)
            blk = re.findall(r"Total blk = (\d+)", out)
            return own, int(blk[-1]) if blk else n
    t_setup, _ = once(1)
    t_run, done = once(frames)
    dt = max(t_run - t_setup, 1e-3)
    return {"frames": done, "threads": threads, "wall_s": t_run, "setup_s": t_setup, "mbps": done * K_INFO / dt / 1e6,
            "note": "reference's own multi-threaded binary (unchanged main + Simulator + thread pool), whole frame incl. "
                    "source/encoder/channel, 'Total time cost' minus a 1-frame run"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    fpp = args.ref_frames
    for _ in range(args.warmup if args.warmup < 2 else 1):  # one untimed pass pages the binary and data files in
        run_reference_sample(max(4, fpp // 8), cores, kind=args.ref_kind)
    vals, secs, last = [], [], None
    for _ in range(args.steps):
        last = run_reference_sample(fpp, cores, kind=args.ref_kind)
        vals.append(last["rx_mbps"])
        secs.append(last["rx_s"])
    v = statistics.mean(vals)
    sample = (f"{args.steps} steps x {cores} processes x {fpp} frames of the workload (LCG-generated, distinct seeds); "
              f"receiver stages only (k-means+resolve+demap+decode) timed inside the reference harness; "
              f"{last['iters_per_frame']:.1f} decoder iterations/frame")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mbit/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": make_config(args),
            "cpu_baseline": {"value": v, "unit": "Mbit/s", "cores": cores, "kind": last["kind"], "sample": sample},
            "e2e": {"value": v, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "iters_per_frame": last["iters_per_frame"], "whole_frame_mbps": last["frame_mbps"],
            "ms_per_frame_per_core": last["ms_per_frame_per_core"]}
    print(json.dumps(line))
    return 0



# ------------------------------------------------------------------------------------------------- extra legs
EXTRA_CONFIGS = [
    # tag, BASELINE config, matrix, modem, 5G, SNR at which the syndrome never clears, frames per step, CPU frames per process, CPU kind
    ("C2-fixedI50", "configs[1]: 5G BG2 R1/2 K960 + 16QAM Gray (4 x 5-iteration metric decodes + final decode; the reference has "
     "no min-sum, so this is its sum-product)", "5GLDPCBG2a3_R12_K960.txt", "4bit_16QAM_Gray.txt", True, -5.0, 8192, 30, "reference"),
    ("C3-fixedI50", "configs[2]: PEG8064 R1/2 + 64QAM Gray", "PEG8064regular0.5.txt", "6bits_64QAM_Gray.txt", False, 0.0, 4096, 12,
     "port"),  # (the reference classes spend 2 x 25 s per process building this code: its C port is timed instead, and says so)
    ("C4-fixedI50", "configs[3]: PEG2304 R1/2 + 16QAM phi1, blind (the reference never resolves it: FER = 1, 50 iterations)",
     "PEG2304regular0.5.txt", "4bit_16QAM_phi1.txt", False, 15.0, 16384, 60, "reference"),
]


def config_leg(kb, torch, dev, stream, smem_peak, spec, steps, with_cpu):
    tag, what, matrix, modem, is_5g, snr_db, B, cpu_frames, cpu_kind = spec
    link = kb.Link(kb.LdpcCode(matrix, is_5g=is_5g), kb.Modem(modem), max_iter=MAX_ITER, kmeans_iter=KMEANS_ITER,
                   early_exit=True, max_batch=B, device=dev.index)
    K, n_tx, E, kw, n_sym = link.code.K, link.code.N_tx, link.code.E, link.k_words, link.n_sym
    var = kb.snr_to_var(snr_db)
    pool = max(2, -(-140 * 2**20 // (B * n_sym * 8)))  # inputs rotate over more than the 126 MiB L2
    ys = [torch.empty((B, n_sym, 2), dtype=torch.float32, device=dev) for _ in range(pool)]
    us = [torch.empty((B, kw), dtype=torch.int32, device=dev) for _ in range(pool)]
    hs = torch.empty((B, 2), dtype=torch.float32, device=dev)
    for i in range(pool):
        link.generate_dev(B, snr_db, 17, i * B, us[i].data_ptr(), hs.data_ptr(), ys[i].data_ptr(), stream)
    uu_hat = torch.empty((B, kw), dtype=torch.int32, device=dev)
    ret = torch.empty((B,), dtype=torch.int32, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)

    def step(i):
        link.receive_dev(B, ys[i % pool].data_ptr(), var, uu_hat.data_ptr(), ret.data_ptr(), stream=stream)
        link.count_errors_dev(B, us[i % pool].data_ptr(), uu_hat.data_ptr(), counters.data_ptr(), stream)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step(3 + i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    iters = float(ret.float().clamp(max=MAX_ITER).mean().item())
    thr = None
    if is_5g:  # BASELINE names "min-sum decoding" for this config; the reference has none → throughput mode, BER/FER-gated only
        thr = {"note": "whole receiver, same step; NOT the reference's algorithm (tests/test_gpu_minsum.py gates them)"}
        for alg, label in ((1, "minsum_flooding_fp32"), (2, "minsum_flooding_fp16x2"), (3, "minsum_layered")):
            link.set_algorithm(alg, 0.8)
            step(0)
            torch.cuda.synchronize()
            a.record()
            for i in range(steps):
                step(3 + i)
            b.record()
            torch.cuda.synchronize()
            thr[label + "_receiver_mbps"] = B * K / (a.elapsed_time(b) / steps * 1e-3) / 1e6
        link.set_algorithm(0)
    hhat = torch.empty((B, 2), dtype=torch.float32, device=dev)
    llr = torch.empty((B, n_tx), dtype=torch.float32, device=dev)
    cc_hat = torch.empty((B, link.words_n), dtype=torch.int32, device=dev)
    link.kmeans_dev(B, ys[0].data_ptr(), hhat.data_ptr(), 0, stream)
    link.demap_dev(B, ys[0].data_ptr(), hhat.data_ptr(), var, llr.data_ptr(), stream)
    link.set_early_exit(False)
    for _ in range(2):
        link.decode_dev(B, llr.data_ptr(), False, MAX_ITER, cc_hat.data_ptr(), ret.data_ptr(), stream)
    a.record()
    for _ in range(3):
        link.decode_dev(B, llr.data_ptr(), False, MAX_ITER, cc_hat.data_ptr(), ret.data_ptr(), stream)
    b.record()
    torch.cuda.synchronize()
    dec_ms = a.elapsed_time(b) / 3
    info = link.decoder_info()
    alg_bytes = 16.0 * E * MAX_ITER * B
    out = {"workload": tag, "what": what, "snr_db": snr_db, "frames_per_step": B, "value": B * K / (ms * 1e-3) / 1e6, "unit": "Mbit/s",
           "ms_per_step": ms, "frames_per_s": B / (ms * 1e-3), "iters_per_frame": iters, "info_bits": K, "edges": E,
           "roofline": {"bound": "smem", "kernel": {0: "bp_regular_kernel<6,3>", 1: "bp_regular_kernel<8,4,1024> (PEG8064)"}.get(
                            info["kernel_kind"], "bp_qc_kernel<BG2>" if info["qc_plan"] else "bp_generic_kernel"),
                        "achieved": alg_bytes / (dec_ms * 1e-3) / 1e9, "peak": smem_peak, "unit": "GB/s",
                        "frac": alg_bytes / (dec_ms * 1e-3) / 1e9 / smem_peak if smem_peak else None, "traffic": None,
                        "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dec_ms,
                        "decode_only_mbps": B * K / (dec_ms * 1e-3) / 1e6}}
    if thr:
        out["throughput_mode"] = thr
    link.close()
    if with_cpu:
        cores = os.cpu_count() or 1
        r = run_reference_sample(cpu_frames, cores, kind=cpu_kind, matrix=matrix, modem=modem, snr_db=snr_db, k_info=K, is_5g=is_5g)
        out["cpu_baseline"] = {"value": r["rx_mbps"], "unit": "Mbit/s", "cores": cores, "kind": r["kind"],
                               "sample": f"{cores} x {cpu_frames} frames, {r['iters_per_frame']:.1f} iterations/frame, wall {r['wall_s']:.1f} s"
                                         + ("" if r["kind"] == "reference" else " (C port of the reference: whole frame, stages not timed apart)")}
    return out


def c5_leg(kb, n_gpus, frames_total=100_000_000, batch=16384):
    """BASELINE config 5 through the product's own multi-GPU driver (kml_sweep_run = Simulator::Simulate): PEG2304 + QPSK,
    0:1:30 dB, 10^8 frames in total, counters of the n_gpus GPUs summed by one ncclAllReduce per SNR point."""
    import hashlib
    import tempfile
    import numpy as np
    per_point = -(-frames_total // 31)
    with tempfile.TemporaryDirectory() as d:
        cfg = os.path.join(d, "config.toml")
        open(cfg, "w").write(
            f"[range]\nminimum_snr = 0.0\nmaximum_snr = 30.0\nstep_snr = 1.0\nmaximum_error_number = 4000000000\n"
            f"maximum_block_number = {per_point}\nthread_block_number = 1000\n[decoder]\ntrue_h_arg = false\n[xcodec]\n5gldpc = false\n"
            f"metric_type = false\nmetric_iter = 5\n[histogram]\nenable = false\n[ldpc]\nmax_iter = {MAX_ITER}\nactive = true\n"
            f"matrix_file = \"{MATRIX}\"\n[modem]\nmodem_file = \"{MODEM}\"\n[gpu]\nseed = 17\ngpus = {n_gpus}\nbatch = {batch}\n")
        sim = kb.Simulator(cfg, data_dir=os.path.join(ROOT, "config"))
        t0 = time.perf_counter()
        snr, ber, fer, cnt = sim.simulate(echo=False)
        wall = time.perf_counter() - t0
        setup_s, sweep_s = sim.last_timing()
    frames = int(cnt[:, 0].sum())
    return {"workload": "C5: PEG2304 R1/2 + QPSK, SNR 0:1:30 dB, blind k-means + BP (early exit, max_iter 50)", "scaling": "strong",
            "n_gpus": n_gpus, "frames": frames, "snr_points": len(snr), "seconds": sweep_s, "setup_seconds": setup_s, "wall_seconds": wall,
            "frames_per_s": frames / sweep_s, "decoded_mbps": frames * K_INFO / sweep_s / 1e6,
            "counters_checksum": hashlib.sha256(np.ascontiguousarray(cnt, np.uint64).tobytes()).hexdigest()[:16],
            "err_blk_total": int(cnt[:, 1].sum()), "err_bit_total": int(cnt[:, 3].sum()),
            "fer_at_0_15_30_db": [float(fer[0]), float(fer[15]), float(fer[30])],
            "reduce": "one ncclAllReduce of 4 x uint64 per SNR point inside kml_sweep_run (single process, one host thread per GPU)",
            "note": "the product's own multi-GPU path; counters_checksum must be identical for every n_gpus (Philox frames derive from the global frame index)"}

# ------------------------------------------------------------------------------------------------- our arm (GPU)
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import kmldpc_b200 as kb
    from kmldpc_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = "unset"
    try:  # host buffers and the copy-issuing thread on the CPUs next to this rank's GPU (matters for e2e at N = 8:
        import pynvml  # eight ranks move 8 x 151 MB per step through host memory)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        numa = f"{len(os.sched_getaffinity(0))} cpus"
    except Exception as e:  # affinity is an optimisation, not a requirement
        numa = f"unavailable ({type(e).__name__})"
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side waiting while rank 0 drives all GPUs in the C5 leg

    B, pool = args.batch, args.pool
    link = kb.Link(kb.LdpcCode(MATRIX), kb.Modem(MODEM), max_iter=MAX_ITER, kmeans_iter=KMEANS_ITER, early_exit=True,
                   max_batch=B, device=local)
    kw = link.k_words
    var = kb.snr_to_var(SNR_DB)
    stream = torch.cuda.current_stream().cuda_stream

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~1 s to come up: start it before the (untimed) input generation
    # ---- synthetic inputs: `pool` batches of B frames, Philox-generated on the device (untimed), > L2 in total
    ys = [torch.empty((B, N_SYM, 2), dtype=torch.float32, device=dev) for _ in range(pool)]
    us = [torch.empty((B, kw), dtype=torch.int32, device=dev) for _ in range(pool)]
    hs = torch.empty((B, 2), dtype=torch.float32, device=dev)
    lo, hi = shard.frame_range(rank, world, world * pool * B)  # this rank's slice of the global frame index space
    for i, (frame0, nfr) in enumerate(shard.batches(lo, hi, B)):
        link.generate_dev(nfr, SNR_DB, 17, frame0, us[i].data_ptr(), hs.data_ptr(), ys[i].data_ptr(), stream)
    uu_hat = torch.empty((B, kw), dtype=torch.int32, device=dev)
    ret = torch.empty((B,), dtype=torch.int32, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step_dev(i):
        link.receive_dev(B, ys[i % pool].data_ptr(), var, uu_hat.data_ptr(), ret.data_ptr(), stream=stream)
        link.count_errors_dev(B, us[i % pool].data_ptr(), uu_hat.data_ptr(), counters.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident
    for i in range(args.warmup):
        step_dev(i)
    counters.zero_()
    barrier()
    sampler.mark_begin()
    l0 = link.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step_dev(args.warmup + i)
    shard.reduce_counters(counters)  # the only collective of the path: 4 x int64 error counters (NCCL)
    ev1.record()
    barrier()
    launches = link.launches - l0
    ms_total = ev0.elapsed_time(ev1)
    iters_mean = float(ret.float().clamp(max=MAX_ITER).mean().item())
    cnt = counters.tolist()

    # ---- e2e: host buffers (pinned), H2D + D2H inside the timed region
    y_host = [torch.empty((B, N_SYM, 2), dtype=torch.float32).pin_memory() for _ in range(min(pool, 3))]
    for i, yh in enumerate(y_host):
        yh.copy_(ys[i])
    uu_host = torch.empty((B, kw), dtype=torch.int32).pin_memory()
    ret_host = torch.empty((B,), dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    uu_host2 = torch.empty((B, kw), dtype=torch.int32).pin_memory()
    ret_host2 = torch.empty((B,), dtype=torch.int32).pin_memory()
    outs = [(uu_host, ret_host), (uu_host2, ret_host2)]

    def step_host(i):
        link.receive_raw(B, y_host[i % len(y_host)].data_ptr(), var, uu_host.data_ptr(), ret_host.data_ptr())

    def submit_host(i):  # pipelined form of the same call: at most one earlier batch still in flight when this one is queued,
        uo, ro = outs[i % 2]  # so its pinned input (3 buffers) and output (2 sets) are free again
        link.receive_submit_raw(B, y_host[i % len(y_host)].data_ptr(), var, uo.data_ptr(), ro.data_ptr())
        link.receive_wait(1)

    for i in range(max(1, min(args.warmup, 3))):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(i)
    torch.cuda.synchronize()
    e2e_block_s = time.perf_counter() - t0
    for i in range(3):
        submit_host(i)
    link.receive_wait(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        submit_host(i)
    link.receive_wait(0)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    step_host(args.steps - 1)  # (uu_host = the last batch, for the comparison below)
    # the same through the reference-typed entry point (std::complex<double> symbols: twice the bytes over PCIe, narrowed
    # to fp32 on the device inside the call) — what a maintainer splicing the call into Simulator::run_blocks would pay
    y64_host = torch.empty((B, N_SYM, 2), dtype=torch.float64).pin_memory()
    y64_host.copy_(ys[0])
    link.receive_f64_raw(B, y64_host.data_ptr(), var, uu_host.data_ptr(), ret_host.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for i in range(max(2, args.steps // 2)):
        link.receive_f64_raw(B, y64_host.data_ptr(), var, uu_host.data_ptr(), ret_host.data_ptr())
    torch.cuda.synchronize()
    e2e64_s = (time.perf_counter() - t0) / max(2, args.steps // 2)
    del y64_host
    step_host(args.steps - 1)  # (uu_host back to the last fp32 batch for the comparison below)
    # decisions of the host path equal those of the device path on the same batch
    link.receive_dev(B, ys[(args.steps - 1) % len(y_host)].data_ptr(), var, uu_hat.data_ptr(), ret.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    same = bool((uu_hat.cpu() == uu_host).all().item())

    # ---- roofline leg: the decoder kernel alone on LLRs resident in HBM
    hhat = torch.empty((B, 2), dtype=torch.float32, device=dev)
    llr = torch.empty((B, N_CODE), dtype=torch.float32, device=dev)
    cc_hat = torch.empty((B, link.words_n), dtype=torch.int32, device=dev)
    link.kmeans_dev(B, ys[0].data_ptr(), hhat.data_ptr(), 0, stream)
    link.demap_dev(B, ys[0].data_ptr(), hhat.data_ptr(), var, llr.data_ptr(), stream)
    link.set_early_exit(False)  # exactly MAX_ITER iterations per frame: what algorithmic_bytes_per_launch assumes
    for _ in range(2):
        link.decode_dev(B, llr.data_ptr(), False, MAX_ITER, cc_hat.data_ptr(), ret.data_ptr(), stream)
    torch.cuda.synchronize()
    reps = max(3, min(10, args.steps))
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(reps):
        link.decode_dev(B, llr.data_ptr(), False, MAX_ITER, cc_hat.data_ptr(), ret.data_ptr(), stream)
    d1.record()
    torch.cuda.synchronize()
    dec_ms = d0.elapsed_time(d1) / reps
    dec_iters = float(MAX_ITER)  # early exit off: every frame executes all of them (decision latched at the first zero syndrome)
    link.set_early_exit(True)
    # the roofline's denominator, measured on this device (rank 0): conflict-free LDS.128 on every SM
    smem_measured = 0.0
    if rank == 0:
        try:
            smem_measured = link.measure_smem_bandwidth()
        except Exception:
            smem_measured = 0.0
    # k-means alone (the metric's second half: k-means frames/s)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    link.kmeans_dev(B, ys[1 % pool].data_ptr(), hhat.data_ptr(), 0, stream)
    k0.record()
    for i in range(reps):
        link.kmeans_dev(B, ys[i % pool].data_ptr(), hhat.data_ptr(), 0, stream)
    k1.record()
    torch.cuda.synchronize()
    km_ms = k0.elapsed_time(k1) / reps
    # … and at config.toml's 15 dB (rank 0): the cell boundaries cross few samples there, so fewer memberships change per pass
    km15_ms = None
    if rank == 0:
        y15 = [torch.empty((B, N_SYM, 2), dtype=torch.float32, device=dev) for _ in range(pool)]
        u_tmp = torch.empty((B, kw), dtype=torch.int32, device=dev)
        for i in range(pool):
            link.generate_dev(B, 15.0, 23, i * B, u_tmp.data_ptr(), hs.data_ptr(), y15[i].data_ptr(), stream)
        link.kmeans_dev(B, y15[0].data_ptr(), hhat.data_ptr(), 0, stream)
        k0.record()
        for i in range(reps):
            link.kmeans_dev(B, y15[i % pool].data_ptr(), hhat.data_ptr(), 0, stream)
        k1.record()
        torch.cuda.synchronize()
        km15_ms = k0.elapsed_time(k1) / reps
        del y15, u_tmp
    sampler.mark_end()

    # ---- secondary: throughput-mode decoders (NOT the reference's algorithm: normalised min-sum, gated by BER/FER tests)
    thr = {}
    if rank == 0 and not args.quick:
        def time_decode(alg, iters):
            link.set_algorithm(alg, 0.8)
            for _ in range(2):
                link.decode_dev(B, llr.data_ptr(), False, iters, cc_hat.data_ptr(), ret.data_ptr(), stream)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                link.decode_dev(B, llr.data_ptr(), False, iters, cc_hat.data_ptr(), ret.data_ptr(), stream)
            b.record()
            torch.cuda.synchronize()
            return B * K_INFO / (a.elapsed_time(b) / 5 * 1e-3) / 1e6
        link.set_early_exit(False)
        for alg, label in ((1, "minsum_fp32"), (2, "minsum_fp16x2")):
            for iters in (50, 10, 5):
                thr[f"{label}_I{iters}_decode_mbps"] = time_decode(alg, iters)
        thr["sum_product_I10_decode_mbps"] = time_decode(0, 10)
        thr["sum_product_I5_decode_mbps"] = time_decode(0, 5)
        link.set_algorithm(2, 0.8)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_dev(0)
        a.record()
        for i in range(5):
            step_dev(i)
        b.record()
        torch.cuda.synchronize()
        thr["minsum_fp16x2_I50_receiver_mbps"] = 5 * B * K_INFO / (a.elapsed_time(b) * 1e-3) / 1e6
        for iters in (10, 5):  # the whole receiver (k-means + resolve + demap + decode) at a smaller fixed iteration count
            l2 = kb.Link(link.code, link.modem, max_iter=iters, kmeans_iter=KMEANS_ITER, early_exit=False, max_batch=B,
                         device=dev.index, algorithm=2)
            for i in range(2):
                l2.receive_dev(B, ys[i % pool].data_ptr(), var, uu_hat.data_ptr(), ret.data_ptr(), stream=stream)
            a.record()
            for i in range(5):
                l2.receive_dev(B, ys[i % pool].data_ptr(), var, uu_hat.data_ptr(), ret.data_ptr(), stream=stream)
            b.record()
            torch.cuda.synchronize()
            thr[f"minsum_fp16x2_I{iters}_receiver_mbps"] = 5 * B * K_INFO / (a.elapsed_time(b) * 1e-3) / 1e6
            l2.close()
        thr["note"] = ("decode-only = kml_decode_dev on HBM-resident LLRs, early exit off, CUDA events; min-sum is not in the "
                       "reference (no parity claim): gated by tests/test_gpu_minsum.py against the sum-product decoder")
        link.set_algorithm(0)
        link.set_early_exit(True)

    # ---- secondary: fused Monte-Carlo path (Philox → … → counters) and the early-exit figure at 15 dB (config.toml SNR)
    t_f = None
    if rank == 0 and not args.quick:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fcnt, fit = link.simulate(SNR_DB, B * 4, seed=99)
        t_f = time.perf_counter() - t0
        t0 = time.perf_counter()
        ecnt, eit = link.simulate(15.0, B * 8, seed=99)
        t_e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- BASELINE configs 2-4 (N = 1 only) and config 5 (every N: rank 0 drives all GPUs of the job, the others wait on the host)
    cfg_legs, c5 = [], None
    if rank == 0 and world == 1 and not args.quick:
        smem_for_legs = smem_measured if smem_measured > 0 else 148 * 128 * 1965.0 * 1e6 / 1e9
        for spec in EXTRA_CONFIGS:
            cfg_legs.append(config_leg(kb, torch, dev, stream, smem_for_legs, spec, 5, not args.no_cpu))
    if not args.no_c5:
        torch.cuda.synchronize()
        if rank == 0:
            c5 = c5_leg(kb, min(world, torch.cuda.device_count()), args.c5_frames, args.batch)
        if cpu_group is not None:
            dist.barrier(group=cpu_group)

    # ---- max over ranks
    times = torch.tensor([ms_total, e2e_s * 1e3, dec_ms, km_ms, e2e64_s * 1e3, e2e_block_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, dec_ms, km_ms, e2e64_ms, e2e_block_ms = times.tolist()
    if rank == 0:
        frames = world * B * args.steps
        value = frames * K_INFO / (ms_total * 1e-3) / 1e6
        e2e = frames * K_INFO / (e2e_ms * 1e-3) / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        smem_derived = n_sm * 128 * sm_max * 1e6 / 1e9        # GB/s: 128 B/clk/SM — not in MEASURED_PEAKS
        smem_peak = smem_measured if smem_measured > 0 else smem_derived
        alg_bytes = 16.0 * N_EDGES * MAX_ITER * B             # 16 B of shared-memory traffic per edge-iteration
        achieved = alg_bytes / (dec_ms * 1e-3) / 1e9
        hbm_bytes = B * (N_CODE * 4 + link.words_n * 4 + 4)  # LLR in + packed decisions + return value out
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic, traffic_src = None, None
        try:  # dram__bytes_read + write of this kernel at this batch size, from this round's `ncu --set full` capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "decoder_traffic.json")))
            if int(tj.get("frames", 0)) == B:
                traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            r = run_reference_sample(args.ref_frames, cores, kind=args.ref_kind)
            cpu = {"value": r["rx_mbps"], "unit": "Mbit/s", "cores": cores, "kind": r["kind"],
                   "sample": f"{cores} processes x {args.ref_frames} frames of the same workload, receiver stages only, "
                             f"{r['iters_per_frame']:.1f} iterations/frame, wall {r['wall_s']:.1f} s",
                   "whole_frame_mbps": r["frame_mbps"], "ms_per_frame_per_core": r["ms_per_frame_per_core"]}
            if not args.quick:
                cpu["stock_binary"] = run_stock_binary(max(cores * 40, 320), cores)
        line = {"metric": METRIC, "value": value, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": make_config(args),
                "frames_per_s": frames / (ms_total * 1e-3), "iters_per_frame": iters_mean,
                "counters": {"tot_blk": cnt[0], "err_blk": cnt[1], "tot_bit": cnt[2], "err_bit": cnt[3]},
                "e2e": {"value": e2e, "unit": "Mbit/s", "h2d_bytes_per_step": B * N_SYM * 8,
                        "d2h_bytes_per_step": B * kw * 4 + B * 4, "ms_per_step": e2e_ms / args.steps,
                        "timer": "host wall clock around `steps` x (kml_receive_submit + kml_receive_wait(1)) + kml_receive_wait(0): every "
                                 "step copies its y from pinned host memory and reads its decisions and return values back",
                        "blocking_call": {"value": frames * K_INFO / (e2e_block_ms * 1e-3) / 1e6, "ms_per_step": e2e_block_ms / args.steps,
                                          "note": "the same batches through the blocking kml_receive, one call per step"},
                        "cpu_affinity": numa,
                        "matches_device_path": same,
                        "reference_types": {"value": world * B * K_INFO / (e2e64_ms * 1e-3) / 1e6, "unit": "Mbit/s",
                                            "h2d_bytes_per_step": B * N_SYM * 16, "ms_per_step": e2e64_ms,
                                            "note": "kml_receive_f64: std::complex<double> symbols from pinned host memory, "
                                                    "narrowed on the device inside the call"}},
                "gpu_launches": int(launches),
                "roofline": {"bound": "smem", "kernel": "bp_regular_kernel<6,3> (BP decoder)", "achieved": achieved,
                             "peak": smem_peak, "unit": "GB/s", "frac": achieved / smem_peak, "traffic": traffic,
                             "traffic_source": traffic_src,
                             "peak_source": (f"measured live on this GPU: conflict-free LDS.128 on all SMs, best of 2 "
                                             f"(kml_measure_smem_bandwidth); derived {n_sm} SMs x 128 B/clk x {sm_max:.0f} MHz "
                                             f"= {smem_derived:.0f} GB/s; shared memory is not in MEASURED_PEAKS.json")
                             if smem_measured > 0 else
                             f"derived: {n_sm} SMs x 128 B/clk x {sm_max:.0f} MHz (shared memory; not in MEASURED_PEAKS.json)",
                             "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dec_ms, "iters_per_frame": dec_iters,
                             "edge_iterations_per_s": N_EDGES * MAX_ITER * B / (dec_ms * 1e-3),
                             "decode_only_mbps": B * K_INFO / (dec_ms * 1e-3) / 1e6,
                             "hbm": {"bound": "hbm", "achieved": hbm_bytes / (dec_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": hbm_bytes / (dec_ms * 1e-3) / 1e9 / hbm_peak,
                                     "note": "LLR in + packed decisions out of the standalone decoder; never binding",
                                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s"}},
                "kmeans": {"frames_per_s": world * B / (km_ms * 1e-3), "ms_per_batch": km_ms, "passes": KMEANS_ITER, "snr_db": SNR_DB,
                           "frames_per_s_per_gpu_at_15dB": B / (km15_ms * 1e-3) if km15_ms else None,
                           "note": "exact-assignment kernel (fp64 estimate and sums, fp32 filter): every frame within 1e-4 of the reference"},
                "clocks": clocks}
        if cpu:
            line["cpu_baseline"] = cpu
        if thr:
            line["throughput_mode"] = thr
        if cfg_legs:
            line["configs"] = cfg_legs
        if c5:
            line["c5"] = c5
        if t_f is not None:
            line["fused_simulate"] = {"mbps": 4 * B * K_INFO / t_f / 1e6, "frames": 4 * B, "iters_per_frame": fit / (4 * B),
                                      "note": "kml_simulate: Philox bits+encode+map+channel+receiver+count, host wall clock"}
            line["early_exit_15dB"] = {"mbps": 8 * B * K_INFO / t_e / 1e6, "frames": 8 * B, "iters_per_frame": eit / (8 * B),
                                       "fer": ecnt[1] / ecnt[0], "ber": ecnt[3] / ecnt[2],
                                       "note": "config.toml's SNR point through kml_simulate with the reference's early exit"}
        print(json.dumps(line))
    link.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="frames per step per GPU")
    ap.add_argument("--pool", type=int, default=2, help="distinct input batches (pool x batch x 9216 B must exceed L2)")
    ap.add_argument("--ref-frames", type=int, default=150, help="frames per host process in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip the secondary fused / early-exit figures and the config 2-4 legs")
    ap.add_argument("--no-c5", action="store_true", help="skip the config-5 sweep leg (10^8 frames through kml_sweep_run)")
    ap.add_argument("--c5-frames", type=int, default=100_000_000, help="total frames of the config-5 sweep")
    ap.add_argument("--ref-kind", default="reference", choices=["reference", "port"],
                    help="CPU arm: the unmodified reference classes (oracle/_ref/ref_harness; missing = error) or, explicitly, their C port")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
