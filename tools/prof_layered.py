"""Throughput of the whole link (kml_simulate) on the 5G BG2 code per decoder algorithm: 0 sum-product (reference), 1 flooding
min-sum, 3 layered min-sum.  Prints frames/s, FER and iterations per frame.  Usage: python tools/prof_layered.py [frames]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from tests import util  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
link = util.gpu_link("5g_16qam_gray_10db", max_batch=16384)
for snr in (8.0, 10.0, 12.0, 15.0):
    for alg in (0, 1, 2, 3):
        link.set_algorithm(alg, 0.8)
        link.simulate(snr, 32768, seed=5)
        t0 = time.perf_counter()
        cnt, it = link.simulate(snr, frames, seed=7)
        dt = time.perf_counter() - t0
        print(f"snr {snr:5.1f} alg {alg}: {frames / dt / 1e3:9.1f} kframes/s  FER {cnt[1] / cnt[0]:.5f}  BER {cnt[3] / cnt[2]:.3e}  "
              f"ret/frame {it / frames:.2f}", flush=True)
link.close()
