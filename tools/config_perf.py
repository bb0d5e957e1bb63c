"""Fused-path throughput of every BASELINE config (kml_simulate, early exit on) — sanity numbers, not bench lines."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmldpc_b200 as kb
CASES = [("C1 PEG2304+QPSK 15dB", "PEG2304regular0.5.txt", "2bits_QPSK.txt", False, 15.0, 65536),
         ("C1 PEG2304+4PSK 6dB", "PEG2304regular0.5.txt", "2bits_4PSK.txt", False, 6.0, 65536),
         ("C2 5G BG2+16QAM 10dB", "5GLDPCBG2a3_R12_K960.txt", "4bit_16QAM_Gray.txt", True, 10.0, 32768),
         ("C3 PEG8064+64QAM 20dB", "PEG8064regular0.5.txt", "6bits_64QAM_Gray.txt", False, 20.0, 16384),
         ("C4 PEG2304+16QAM phi1 15dB", "PEG2304regular0.5.txt", "4bit_16QAM_phi1.txt", False, 15.0, 32768),
         ("C4 PEG2304+16QAM Gray 12dB", "PEG2304regular0.5.txt", "4bit_16QAM_Gray.txt", False, 12.0, 65536)]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, mat, mod, g5, snr, frames in CASES:
    if only and only not in name: continue
    link = kb.Link(kb.LdpcCode(mat, is_5g=g5), kb.Modem(mod), max_batch=8192)
    link.simulate(snr, 8192, seed=1)
    t0 = time.perf_counter(); cnt, it = link.simulate(snr, frames, seed=2); dt = time.perf_counter() - t0
    print(f"{name:32s} {frames/dt/1e3:9.1f} kframes/s {frames*link.code.K/dt/1e6:9.1f} Mbit/s  iters/frame {it/frames:5.1f}  FER {cnt[1]/cnt[0]:.3f} BER {cnt[3]/cnt[2]:.4f}")
    link.close()
