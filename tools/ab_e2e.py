"""Pipelined e2e (kml_receive_submit / wait) of the headline workload, for A/B runs of the sub-batch size (KML_TUNING build)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmldpc_b200 as kb
B, steps = 16384, 12
link = kb.Link(kb.LdpcCode("PEG2304regular0.5.txt"), kb.Modem("2bits_QPSK.txt"), max_batch=B)
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
kw = link.k_words
ys = [torch.empty((B, 1152, 2), dtype=torch.float32, device=dev) for _ in range(3)]
us = torch.empty((B, kw), dtype=torch.int32, device=dev); hs = torch.empty((B, 2), dtype=torch.float32, device=dev)
for i in range(3):
    link.generate_dev(B, -5.0, 17, i * B, us.data_ptr(), hs.data_ptr(), ys[i].data_ptr(), st)
torch.cuda.synchronize()
yh = [y.cpu().pin_memory() for y in ys]
outs = [(torch.empty((B, kw), dtype=torch.int32).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory()) for _ in range(2)]
var = kb.snr_to_var(-5.0)
def run(n):
    for i in range(n):
        link.receive_submit_raw(B, yh[i % 3].data_ptr(), var, outs[i % 2][0].data_ptr(), outs[i % 2][1].data_ptr())
        link.receive_wait(1)
    link.receive_wait(0)
run(3)
t0 = time.perf_counter(); run(steps); dt = time.perf_counter() - t0
print(f"chunk={os.environ.get('KML_RX_CHUNK','default')} slow={os.environ.get('KML_RX_SLOW','default')}: {dt / steps * 1e3:.3f} ms/step = {B * 1152 / (dt / steps) / 1e6:.0f} Mbit/s")
