"""One fused Monte-Carlo pass (kml_simulate, early exit on) of a named configuration — the command the ncu launch lists
and front-end captures of the early-exit regime are taken from.  usage: prof_frontend.py <case> <snr_db> <frames> [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmldpc_b200 as kb

CASES = {"C1q": ("PEG2304regular0.5.txt", "2bits_QPSK.txt", False), "C1p": ("PEG2304regular0.5.txt", "2bits_4PSK.txt", False),
         "C2": ("5GLDPCBG2a3_R12_K960.txt", "4bit_16QAM_Gray.txt", True), "C3": ("PEG8064regular0.5.txt", "6bits_64QAM_Gray.txt", False),
         "C4g": ("PEG2304regular0.5.txt", "4bit_16QAM_Gray.txt", False), "C4p": ("PEG2304regular0.5.txt", "4bit_16QAM_phi1.txt", False)}
case, snr, frames = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
mat, mod, g5 = CASES[case]
link = kb.Link(kb.LdpcCode(mat, is_5g=g5), kb.Modem(mod), max_batch=batch)
link.simulate(snr, batch, seed=1)
t0 = time.perf_counter()
cnt, it = link.simulate(snr, frames, seed=2)
dt = time.perf_counter() - t0
print(f"{case} {snr} dB: {frames / dt / 1e6:.2f} M frames/s, {frames * link.code.K / dt / 1e6:.0f} Mbit/s, "
      f"iters/frame {it / frames:.2f}, FER {cnt[1] / cnt[0]:.4f}, launches {link.launches}")
link.close()
