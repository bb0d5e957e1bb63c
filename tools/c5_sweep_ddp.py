"""BASELINE config C5 as north_star states it: PEG2304 + QPSK/4PSK, SNR 0:1:30 dB, 10^8 frames in total, frames sharded
over the ranks of a torchrun job (one process per GPU), the 4 error counters of every SNR point all-reduced ONCE over NCCL.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_sweep_ddp.py
Device time = max over ranks (CUDA events around each rank's share); counters are independent of the number of ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import kmldpc_b200 as kb
from kmldpc_b200 import shard

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
modem = os.environ.get("MODEM", "2bits_QPSK.txt")
total = int(float(os.environ.get("TOTAL", 1e8)))
snrs = np.arange(0.0, 30.5, 1.0)
per_point = total // len(snrs)
link = kb.Link(kb.LdpcCode("PEG2304regular0.5.txt"), kb.Modem(modem), max_iter=50, max_batch=16384, device=local)
link.simulate(10.0, 16384, seed=1)  # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
cnt = np.zeros((len(snrs), 4), np.int64)
for i, s in enumerate(snrs):
    lo, hi = shard.frame_range(rank, world, per_point)
    c, _ = link.simulate(float(s), hi - lo, seed=17 + i, frame_begin=lo, max_err_blk=0)
    cnt[i] = c.astype(np.int64)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
tc = torch.from_numpy(cnt).cuda()
if world > 1:
    shard.reduce_counters(tc)                      # the ONE collective of the path: 31 x 4 x int64 over NCCL
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
wall = time.time() - t0
if rank == 0:
    c = tc.cpu().numpy()
    frames = int(c[:, 0].sum())
    sec = ms.item() / 1e3
    print(f"C5 {modem} on {world} GPU(s): {frames} frames, {len(snrs)} SNR points, device time (max over ranks) {sec:.2f} s "
          f"= {frames / sec / 1e6:.2f} M frames/s = {frames * 1152 / sec / 1e9:.2f} Gbit/s decoded (wall {wall:.2f} s)")
    print("counters checksum", int(c[:, 1].sum()), int(c[:, 3].sum()))
    for i in range(0, len(snrs), 5):
        print(f"  {snrs[i]:5.1f} dB  BER {c[i, 3] / c[i, 2]:.6f}  FER {c[i, 1] / c[i, 0]:.6f}")
link.close()
if world > 1:
    dist.destroy_process_group()
