"""The C5 leg alone (10^8 frames, 31 SNR points through kml_sweep_run) on the first N GPUs.  Usage: python tools/c5_only.py N [N ...]"""
import json
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
import kmldpc_b200 as kb  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [1]:
    r = bench.c5_leg(kb, n)
    print(json.dumps({k: r[k] for k in ("n_gpus", "seconds", "setup_seconds", "frames_per_s", "counters_checksum")}), flush=True)
