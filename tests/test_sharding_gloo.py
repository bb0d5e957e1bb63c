"""N > 1 host logic on CPU: world_size-2 gloo processes shard a frame range, each produces counters for its frames, the
all-reduce gives the single-process totals (the GPU path does the same with NCCL; frames are pure functions of their
global index)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kmldpc_b200 import shard


def fake_frame_errors(frame: int, k: int = 1152) -> int:
    """Stand-in for a decoded frame: deterministic in the GLOBAL frame index, like the Philox-generated frames."""
    x = (frame * 2654435761) & 0xFFFFFFFF
    return 0 if x % 3 else (x >> 7) % k


def local_counters(lo, hi, batch, k=1152):
    c = torch.zeros(4, dtype=torch.int64)
    for f0, n in shard.batches(lo, hi, batch):
        errs = [fake_frame_errors(f) for f in range(f0, f0 + n)]
        c += torch.tensor([n, sum(e > 0 for e in errs), n * k, sum(errs)], dtype=torch.int64)
    return c


def _worker(rank, world, port, total, batch, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.frame_range(rank, world, total, begin=1000)
    c = local_counters(lo, hi, batch)
    shard.reduce_counters(c)
    if rank == 0:
        np.save(out, c.numpy())
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_frame_range_partition_properties():
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 7, 64, 1000, 100003):
            rs = [shard.frame_range(r, world, total, begin=5) for r in range(world)]
            assert rs[0][0] == 5 and rs[-1][1] == 5 + total
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.frame_range(2, 2, 10)
    assert list(shard.batches(3, 10, 4)) == [(3, 4), (7, 3)]
    assert shard.ber_fer([10, 2, 11520, 57]) == (57 / 11520, 0.2)


@pytest.mark.parametrize("world,total,batch", [(2, 1001, 64), (2, 37, 16)])
def test_two_rank_gloo_reduction_equals_single_process(tmp_path, world, total, batch):
    out = str(tmp_path / "c.npy")
    mp.spawn(_worker, args=(world, _free_port(), total, batch, out), nprocs=world, join=True)
    got = np.load(out)
    want = local_counters(1000, 1000 + total, 97).numpy()  # a different batch size: counts are split-invariant
    assert np.array_equal(got, want)
    assert got[0] == total and got[2] == total * 1152
