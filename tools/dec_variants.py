"""Times the decoder micro-variants (KML_DEC_VARIANT) and checks each against variant 0 on the same LLRs."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, numpy as np
    import kmldpc_b200 as kb
    B = 16384; snr = float(sys.argv[2])
    link = kb.Link(kb.LdpcCode("PEG2304regular0.5.txt"), kb.Modem("2bits_4PSK.txt"), max_iter=50, max_batch=B)
    dev = torch.device("cuda", 0); s = torch.cuda.current_stream().cuda_stream
    y = torch.empty((B, link.n_sym, 2), dtype=torch.float32, device=dev); u = torch.empty((B, link.k_words), dtype=torch.int32, device=dev)
    h = torch.empty((B, 2), dtype=torch.float32, device=dev); llr = torch.empty((B, 2304), dtype=torch.float32, device=dev)
    cc = torch.empty((B, link.words_n), dtype=torch.int32, device=dev); ret = torch.empty((B,), dtype=torch.int32, device=dev)
    link.generate_dev(B, snr, 17, 0, u.data_ptr(), h.data_ptr(), y.data_ptr(), s)
    link.demap_dev(B, y.data_ptr(), h.data_ptr(), kb.snr_to_var(snr), llr.data_ptr(), s)   # true h
    for _ in range(2): link.decode_dev(B, llr.data_ptr(), False, 50, cc.data_ptr(), ret.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): link.decode_dev(B, llr.data_ptr(), False, 50, cc.data_ptr(), ret.data_ptr(), s)
    e1.record(); torch.cuda.synchronize()
    np.save(sys.argv[3], np.concatenate([ret.cpu().numpy()[:, None], cc.cpu().numpy()], axis=1))
    print(json.dumps({"ms": e0.elapsed_time(e1) / 5, "iters": float(ret.float().clamp(max=50).mean())}))
    sys.exit(0)
import numpy as np
for snr in (-5.0, 4.0):
    base = None
    for v in range(8):
        env = dict(os.environ, KML_DEC_VARIANT=str(v))
        out = subprocess.run([sys.executable, __file__, "child", str(snr), f"/tmp/v{v}.npy"], env=env, capture_output=True, text=True)
        try: r = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception: print("variant", v, "failed", out.stderr[-300:]); continue
        a = np.load(f"/tmp/v{v}.npy")
        if base is None: base = a
        ret_same = (a[:, 0] == base[:, 0]).mean(); conv = base[:, 0] < 50
        bits_same = (a[conv, 1:] == base[conv, 1:]).all(axis=1).mean() if conv.any() else 1.0
        print(f"snr {snr:5.1f} variant {v}: {r['ms']:.3f} ms  iters {r['iters']:.2f}  ret identical {ret_same:.5f}  converged-frame bits identical {bits_same:.5f} (n_conv {conv.sum()})")
