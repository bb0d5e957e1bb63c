"""k-means alone (kml_kmeans_dev on HBM-resident symbols), CUDA events.  usage: prof_kmeans.py <snr_db> [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmldpc_b200 as kb
snr = float(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
link = kb.Link(kb.LdpcCode("PEG2304regular0.5.txt"), kb.Modem("2bits_QPSK.txt"), max_batch=B)
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
ys = [torch.empty((B, 1152, 2), dtype=torch.float32, device=dev) for _ in range(2)]
us = torch.empty((B, link.k_words), dtype=torch.int32, device=dev); hs = torch.empty((B, 2), dtype=torch.float32, device=dev)
for i in range(2):
    link.generate_dev(B, snr, 17, i * B, us.data_ptr(), hs.data_ptr(), ys[i].data_ptr(), st)
hh = torch.empty((B, 2), dtype=torch.float32, device=dev)
for i in range(3):
    link.kmeans_dev(B, ys[i % 2].data_ptr(), hh.data_ptr(), 0, st)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(10):
    link.kmeans_dev(B, ys[i % 2].data_ptr(), hh.data_ptr(), 0, st)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"k-means {snr} dB: {ms * 1e3:.1f} us per {B} frames = {B / ms / 1e3:.1f} M frames/s")
