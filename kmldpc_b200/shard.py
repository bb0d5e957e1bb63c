"""Multi-GPU plumbing of the path: frames are independent, so ranks take disjoint ranges of the global
frame index space (Philox counters derive from the global index → any split generates the same frames) and the only
collective is a sum of the 4 error counters.  One process per GPU; torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations


def frame_range(rank: int, world: int, total: int, begin: int = 0) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of frames `begin .. begin+total` owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    base, extra = divmod(total, world)
    lo = begin + rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def batches(lo: int, hi: int, batch: int):
    """Yield (frame0, count) chunks of at most `batch` frames covering [lo, hi)."""
    f = lo
    while f < hi:
        n = min(batch, hi - f)
        yield f, n
        f += n


def reduce_counters(counters):
    """Sum a tensor of counters (tot_blk, err_blk, tot_bit, err_bit[, …]) over all ranks, in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def ber_fer(counters) -> tuple[float, float]:
    """SourceSink::ber / fer (lib/lab/src/sourcesink.cc:43-44) from 64-bit counters."""
    tot_blk, err_blk, tot_bit, err_bit = [int(x) for x in counters[:4]]
    return (err_bit / tot_bit if tot_bit else 0.0, err_blk / tot_blk if tot_blk else 0.0)


class CounterComm:
    """kml_comm_init / kml_reduce_counters: the single-process flavour of the same reduction (GPUs 0 .. n-1 of this
    process, ncclCommInitAll, ONE ncclAllReduce of uint64 words over NVLink) — what kml_sweep_run uses per SNR point."""

    def __init__(self, n_gpus: int):
        import ctypes as C
        from . import capi
        self._lib = capi.load()
        self._h = C.c_void_p()
        rc = self._lib.kml_comm_init(int(n_gpus), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"kml_comm_init: {self._lib.kml_last_error(None).decode()} (rc={rc})")
        self.n_gpus = int(n_gpus)

    def reduce(self, per_gpu):
        """per_gpu: uint64 [n_gpus, count] → their sum, uint64 [count]."""
        import numpy as np
        from . import capi
        a = np.ascontiguousarray(per_gpu, np.uint64).reshape(self.n_gpus, -1)
        out = np.zeros(a.shape[1], np.uint64)
        rc = self._lib.kml_reduce_counters(self._h, a.ctypes.data_as(capi.c_u64p), a.shape[1], out.ctypes.data_as(capi.c_u64p))
        if rc != 0:
            raise RuntimeError(f"kml_reduce_counters: {self._lib.kml_last_error(None).decode()} (rc={rc})")
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self._lib.kml_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
