"""Small invocations of every kernel family for compute-sanitizer --tool memcheck (tiny batches: the tool is ~50x slower)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kmldpc_b200 as kb
CASES = [("PEG2304regular0.5.txt", "2bits_QPSK.txt", False, {}, 15.0), ("PEG2304regular0.5.txt", "4bit_16QAM_Gray.txt", False, {}, 12.0),
         ("PEG2304regular0.5.txt", "2bits_4PSK.txt", False, {"metric_type": True}, 18.0),
         ("5GLDPCBG2a3_R12_K960.txt", "4bit_16QAM_Gray.txt", True, {}, 10.0), ("PEG8064regular0.5.txt", "6bits_64QAM_Gray.txt", False, {}, 20.0)]
for mat, mod, g5, opts, snr in CASES:
    link = kb.Link(kb.LdpcCode(mat, is_5g=g5), kb.Modem(mod), max_batch=24, **opts)
    cnt, it = link.simulate(snr, 60, seed=3)                      # 24 + 24 + 12: ragged last batch, both lanes
    u, c, h, y = link.generate(37, snr, seed=4)
    uu, hh, ks, ret, met = link.receive(y, 10 ** (-snr / 10), with_metric=True)
    uu64, h64, _, ret64 = link.receive_f64(y.astype(np.complex128), 10 ** (-snr / 10))
    m2, c2 = link.histogram(snr, 30, seed=5)
    hm = link.histogram_rx(y, 10 ** (-snr / 10), kb.pack_bits(u))
    llr = link.demap(y, hh, 10 ** (-snr / 10))
    link.decode(llr[:5])
    link.decode_p0(1 / (1 + np.exp(-llr[:5].astype(np.float64))))
    if not opts.get("metric_type"):
        for alg in (1, 2):
            link.set_algorithm(alg, 0.8)
            link.decode(llr[:7])
        link.set_algorithm(0)
    print(mat, mod, opts, "ok", cnt[:2], int(ret.sum()), np.array_equal(ret, ret64))
    link.close()
