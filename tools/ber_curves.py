"""BER/FER curves: CUDA path (Philox frames, many) vs the CPU oracle (= the reference bit for bit; LCG frames, fewer) with
95 % Wilson intervals on the oracle's FER.  Writes a table for profiles/."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kmldpc_b200 as kb
from oracle import kml_oracle as ko
from tests.util import wilson

CASES = [("PEG2304 + 4bit_16QAM_Gray (blind)", "PEG2304regular0.5.txt", "4bit_16QAM_Gray.txt", False, False, [5, 10, 15, 20, 25, 30]),
         ("PEG2304 + 2bits_QPSK (blind; 0/180 tie floor)", "PEG2304regular0.5.txt", "2bits_QPSK.txt", False, False, [0, 5, 10, 15, 20]),
         ("PEG2304 + 2bits_4PSK (blind)", "PEG2304regular0.5.txt", "2bits_4PSK.txt", False, False, [0, 5, 10, 15, 20]),
         ("PEG2304 + 4bit_16QAM_phi1 (known h)", "PEG2304regular0.5.txt", "4bit_16QAM_phi1.txt", False, True, [10, 15, 20]),
         ("5G BG2 K960 + 4bit_16QAM_Gray (blind, metric = 4 x 5 iterations)", "5GLDPCBG2a3_R12_K960.txt", "4bit_16QAM_Gray.txt", True, False, [6, 10, 14, 18]),
         ("PEG8064 + 6bits_64QAM_Gray (blind)", "PEG8064regular0.5.txt", "6bits_64QAM_Gray.txt", False, False, [15, 20, 25, 30])]
only = os.environ.get("ONLY")
if only:
    CASES = [c for c in CASES if only in c[0]]
GPU_FRAMES, REF_FRAMES = int(os.environ.get("GPU_FRAMES", 200000)), int(os.environ.get("REF_FRAMES", 6000))
out = []
for name, mat, mod, g5, known, snrs in CASES:
    link = kb.Link(kb.LdpcCode(mat, is_5g=g5), kb.Modem(mod), known_h=known, max_batch=8192)
    olink = ko.Link(mat, mod, is_5g=g5, known_h=known)
    out.append(f"## {name}: CUDA {GPU_FRAMES} Philox frames per point | oracle {REF_FRAMES} LCG frames per point")
    out.append("snr_dB  fer_gpu    ber_gpu     fer_ref  [95% Wilson]        ber_ref    inside")
    for pi, snr in enumerate(snrs):
        cnt, _ = link.simulate(float(snr), GPU_FRAMES, seed=101 + pi)
        # a different block of the per-frame LCG seed space for every point: the same seeds would reuse the same fades
        ref = olink.bulk(float(snr), REF_FRAMES, frame0=(len(out) * 131 + pi) * REF_FRAMES)
        fe = int((ref["nerr"] > 0).sum()); lo, hi = wilson(fe, REF_FRAMES)
        fer_g, ber_g = cnt[1] / cnt[0], cnt[3] / cnt[2]
        ber_r = ref["nerr"].sum() / (REF_FRAMES * olink.code.K)
        out.append(f"{snr:5.1f}  {fer_g:.5f}  {ber_g:.6f}   {fe / REF_FRAMES:.5f}  [{lo:.5f}, {hi:.5f}]  {ber_r:.6f}   {'yes' if lo <= fer_g <= hi else 'NO'}")
    link.close()
print("\n".join(out))
