"""Small driver for ncu: a few standalone launches of the BP decoder (and the receiver chain) on Philox frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmldpc_b200 as kb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
snr = float(sys.argv[3]) if len(sys.argv) > 3 else -5.0
matrix = sys.argv[4] if len(sys.argv) > 4 else "PEG2304regular0.5.txt"
modem = sys.argv[5] if len(sys.argv) > 5 else "2bits_QPSK.txt"
code, mod = kb.LdpcCode(matrix, is_5g=matrix.startswith("5G")), kb.Modem(modem)
link = kb.Link(code, mod, max_iter=50, max_batch=B, algorithm=int(os.environ.get('KML_ALG', '0')))
dev = torch.device("cuda", 0)
s = torch.cuda.current_stream().cuda_stream
y = torch.empty((B, link.n_sym, 2), dtype=torch.float32, device=dev)
u = torch.empty((B, link.k_words), dtype=torch.int32, device=dev)
h = torch.empty((B, 2), dtype=torch.float32, device=dev)
hhat = torch.empty((B, 2), dtype=torch.float32, device=dev)
llr = torch.empty((B, code.N_tx), dtype=torch.float32, device=dev)
cc = torch.empty((B, link.words_n), dtype=torch.int32, device=dev)
ret = torch.empty((B,), dtype=torch.int32, device=dev)
uu = torch.empty((B, link.k_words), dtype=torch.int32, device=dev)
link.generate_dev(B, snr, 17, 0, u.data_ptr(), h.data_ptr(), y.data_ptr(), s)
link.kmeans_dev(B, y.data_ptr(), hhat.data_ptr(), 0, s)
link.demap_dev(B, y.data_ptr(), hhat.data_ptr(), kb.snr_to_var(snr), llr.data_ptr(), s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
link.decode_dev(B, llr.data_ptr(), False, 50, cc.data_ptr(), ret.data_ptr(), s)
e0.record()
for _ in range(reps):
    link.decode_dev(B, llr.data_ptr(), False, 50, cc.data_ptr(), ret.data_ptr(), s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"decode B={B} {ms:.3f} ms/launch  {B/ms*1e3:.0f} frames/s  {B*code.K/ms/1e3:.1f} Mbit/s  iters {ret.float().clamp(max=50).mean().item():.2f}")
link.receive_dev(B, y.data_ptr(), kb.snr_to_var(snr), uu.data_ptr(), ret.data_ptr(), stream=s)
torch.cuda.synchronize()
link.close()
