"""Runs the same sweep on 1 and on all visible GPUs through kml_sweep_run and checks the counters are identical."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import kmldpc_b200 as kb
from kmldpc_b200.link import CONFIG_DIR
G = torch.cuda.device_count()
cfg = open(os.path.join(CONFIG_DIR, "config.toml")).read()
cfg = cfg.replace("maximum_error_number = 1", "maximum_error_number = 100000000").replace("maximum_block_number = 1", "maximum_block_number = 60000")
cfg = cfg.replace("minimum_snr = 15.0", "minimum_snr = 5.0").replace("4bit_16QAM_Gray.txt", "2bits_4PSK.txt")
res = {}
for g in (1, G):
    with tempfile.NamedTemporaryFile("w", suffix=".toml", delete=False) as f:
        f.write(cfg + f"\n[gpu]\nseed = 17\ngpus = {g}\nbatch = 4096\n")
    sim = kb.Simulator(f.name, data_dir=CONFIG_DIR)
    t0 = time.time(); snr, ber, fer, cnt = sim.simulate(echo=False); dt = time.time() - t0
    res[g] = cnt; print(f"gpus={g}: {dt:.2f}s  points {list(snr)}  FER {fer}  counters {cnt.tolist()}")
print("identical counters:", np.array_equal(res[1], res[G]))
