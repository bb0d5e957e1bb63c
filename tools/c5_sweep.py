"""BASELINE config C5: PEG2304 + QPSK/4PSK, SNR 0:1:30 dB, 10^8 frames in total, through kml_sweep_run on all visible GPUs."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import kmldpc_b200 as kb
from kmldpc_b200.link import CONFIG_DIR
G = int(os.environ.get("GPUS", torch.cuda.device_count()))
modem = os.environ.get("MODEM", "2bits_QPSK.txt")
total = int(float(os.environ.get("TOTAL", 1e8)))
per_point = total // 31
cfg = open(os.path.join(CONFIG_DIR, "config.toml")).read()
cfg = cfg.replace("minimum_snr = 15.0", "minimum_snr = 0.0").replace("maximum_snr = 15.0", "maximum_snr = 30.0").replace("step_snr = 5.0", "step_snr = 1.0")
cfg = cfg.replace("maximum_error_number = 1", "maximum_error_number = 2000000000").replace("maximum_block_number = 1", f"maximum_block_number = {per_point}")
cfg = cfg.replace("4bit_16QAM_Gray.txt", modem)
with tempfile.NamedTemporaryFile("w", suffix=".toml", delete=False) as f:
    f.write(cfg + f"\n[gpu]\nseed = 17\ngpus = {G}\nbatch = 16384\n")
sim = kb.Simulator(f.name, data_dir=CONFIG_DIR)
t0 = time.time(); snr, ber, fer, cnt = sim.simulate(echo=False); dt = time.time() - t0
frames = int(cnt[:, 0].sum())
print(f"C5 {modem} on {G} GPU(s): {frames} frames, {len(snr)} SNR points in {dt:.2f} s wall (incl. setup) = {frames / dt / 1e6:.2f} M frames/s = {frames * 1152 / dt / 1e9:.2f} Gbit/s decoded")
print("counters checksum", int(cnt[:, 1].sum()), int(cnt[:, 3].sum()))
for s, b, f_ in zip(snr[::5], ber[::5], fer[::5]): print(f"  {s:5.1f} dB  BER {b:.6f}  FER {f_:.6f}")
