"""Throughput-mode decoder (algorithm = 1, normalised min-sum).  NOT reference-pinned — the reference has no min-sum
(SURVEY §0.3).  Two gates instead: (1) the CUDA kernel against an independent numpy statement of the same algorithm
(oracle/minsum_ref.py) on the reference's own frames; (2) BER/FER against the sum-product decoder (the reference's
algorithm) on the same Philox frames, inside the binomial interval plus the small known min-sum loss."""
import numpy as np
import pytest

from oracle import minsum_ref
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,frames", [("peg2304_4psk_6db", 120), ("peg2304_16qam_gray_12db", 120),
                                         ("5g_16qam_gray_10db", 60), ("peg8064_64qam_20db", 12)])
def test_minsum_kernel_matches_numpy_statement(name, frames):
    olink, rs = util.oracle_frames(name, frames)
    link = util.gpu_link(name, algorithm=1)
    ex = olink.code.export(with_enc=False)
    llr = np.stack([util.llr_of_p0(r.p0) for r in rs]).astype(np.float32)
    cc, uu, ret = link.decode(llr)
    rret, rcc = minsum_ref.decode(ex["row_ptr"], ex["col_idx"], olink.code.N, olink.code.two_z, llr, 50, 50, alpha=0.8)
    # fp32 summation order differs (gather order vs CSR order) and the kernel borrows the mantissa LSB: allow a few
    # frames to leave one iteration apart, demand identical words wherever both stop at the same iteration
    assert (ret == rret).mean() >= 0.95, np.where(ret != rret)[0][:8]
    both = (ret == rret) & (rret < 50)
    assert both.sum() > 0 or "8064" in name
    assert np.array_equal(cc[both], rcc[both])
    link.close()


@pytest.mark.parametrize("name,snr,frames", [("peg2304_4psk_6db", 6.0, 30000), ("peg2304_4psk_6db", 12.0, 30000),
                                             ("peg2304_16qam_gray_12db", 15.0, 30000)])
def test_minsum_ber_fer_gate_against_sum_product(name, snr, frames):
    link = util.gpu_link(name, max_batch=8192)
    spa, it_spa = link.simulate(snr, frames, seed=23)
    link.set_algorithm(1, 0.8)
    ms, it_ms = link.simulate(snr, frames, seed=23)          # the same frames
    link.set_algorithm(0)
    again, _ = link.simulate(snr, frames, seed=23)
    assert np.array_equal(again, spa)                         # switching back restores the reference decoder exactly
    fer_spa, fer_ms = spa[1] / spa[0], ms[1] / ms[0]
    sd = np.sqrt(fer_spa * (1 - fer_spa) / frames)
    # same frames → the difference is far below the binomial spread; min-sum may lose a little, never gain much
    assert fer_ms <= fer_spa + 3 * sd + 0.01, (fer_ms, fer_spa)
    assert fer_ms >= fer_spa - 3 * sd - 0.005, (fer_ms, fer_spa)
    ber_spa, ber_ms = spa[3] / spa[2], ms[3] / ms[2]
    assert abs(ber_ms - ber_spa) <= 0.1 * ber_spa + 2e-3, (ber_ms, ber_spa)
    assert it_ms <= 1.3 * it_spa + frames                     # converges in a comparable number of iterations
    link.close()


@pytest.mark.parametrize("alg", [1, 2])
def test_offset_minsum_variant(alg):
    """|c2v| = max(alpha m - beta, 0) with alpha = 1, beta = 0.5 (offset min-sum): the fp32 kernel against the numpy statement,
    and both message formats against the sum-product decoder on the same frames (BER/FER gate)."""
    name, snr, frames = "peg2304_4psk_6db", 6.0, 30000
    link = util.gpu_link(name, max_batch=8192)
    spa, _ = link.simulate(snr, frames, seed=41)
    link.set_algorithm(alg, 0.8)
    link.set_minsum(1.0, 0.5)
    ms, _ = link.simulate(snr, frames, seed=41)
    fer_spa, fer_ms = spa[1] / spa[0], ms[1] / ms[0]
    sd = np.sqrt(fer_spa * (1 - fer_spa) / frames)
    assert fer_spa - 3 * sd - 0.005 <= fer_ms <= fer_spa + 3 * sd + 0.015, (fer_ms, fer_spa)
    if alg == 1:
        olink, rs = util.oracle_frames(name, 120)
        ex = olink.code.export(with_enc=False)
        llr = np.stack([util.llr_of_p0(r.p0) for r in rs]).astype(np.float32)
        cc, uu, ret = link.decode(llr)
        rret, rcc = minsum_ref.decode(ex["row_ptr"], ex["col_idx"], olink.code.N, olink.code.two_z, llr, 50, 50, alpha=1.0, beta=0.5)
        assert (ret == rret).mean() >= 0.95
        both = (ret == rret) & (rret < 50)
        assert both.sum() > 0 and np.array_equal(cc[both], rcc[both])
    link.close()


def test_minsum_rejects_soft_metric():
    import kmldpc_b200 as kb
    code, mod = kb.LdpcCode("PEG2304regular0.5.txt"), kb.Modem("2bits_4PSK.txt")
    with pytest.raises(kb.KmlError, match="sum-product"):
        kb.Link(code, mod, metric_type=True, algorithm=1)
    link = kb.Link(code, mod, metric_type=True)
    with pytest.raises(kb.KmlError, match="sum-product"):
        link.set_algorithm(1)
    link.close()


@pytest.mark.parametrize("name,snr,frames", [("peg2304_4psk_6db", 6.0, 30000), ("peg2304_16qam_gray_12db", 15.0, 30000),
                                             ("peg8064_64qam_20db", 22.0, 6000), ("5g_16qam_gray_10db", 10.0, 20000),
                                             ("5g_16qam_gray_10db", 14.0, 20000)])
def test_fp16_two_frame_minsum_gate(name, snr, frames):
    """algorithm = 2 (fp16 messages, two frames per shared-memory word): BER/FER gate against the sum-product decoder on the
    same frames, and pairing must not couple frames: odd batches and shifted pairings give the same per-frame answers."""
    link = util.gpu_link(name, max_batch=4096)
    spa, _ = link.simulate(snr, frames, seed=29)
    link.set_algorithm(2, 0.8)
    ms, it_ms = link.simulate(snr, frames, seed=29)
    fer_spa, fer_ms = spa[1] / spa[0], ms[1] / ms[0]
    sd = np.sqrt(max(fer_spa * (1 - fer_spa), 1e-4) / frames)
    assert fer_ms <= fer_spa + 3 * sd + 0.01 and fer_ms >= fer_spa - 3 * sd - 0.005, (fer_ms, fer_spa)
    assert abs(ms[3] / ms[2] - spa[3] / spa[2]) <= 0.1 * spa[3] / spa[2] + 2e-3
    # pairing independence
    u, c, h, y = link.generate(65, snr, seed=31)
    llr = link.demap(y, h, 10 ** (-0.1 * snr))
    cc_all, _, ret_all = link.decode(llr)              # pairs (0,1), (2,3) … (64, 64)
    cc_odd, _, ret_odd = link.decode(llr[1:])          # pairs (1,2), (3,4) …
    assert np.array_equal(cc_all[1:], cc_odd) and np.array_equal(ret_all[1:], ret_odd)
    cc_one, _, ret_one = link.decode(llr[64:65])
    assert np.array_equal(cc_one[0], cc_all[64]) and ret_one[0] == ret_all[64]
    # fixed-iteration mode latches per frame
    link.set_early_exit(False)
    cc_fix, _, ret_fix = link.decode(llr)
    assert np.array_equal(ret_fix, ret_all) and np.array_equal(cc_fix[ret_all < 50], cc_all[ret_all < 50])
    link.close()


def test_layered_kernel_matches_numpy_statement():
    """algorithm = 3 (layered min-sum on the quasi-cyclic structure, bp_layered.cu) against oracle/minsum_ref.decode_layered
    on the reference's own 5G frames: one conflict-free order of updates, so return values and words agree exactly."""
    name = "5g_16qam_gray_10db"
    olink, rs = util.oracle_frames(name, 61)              # odd count: the last CTA holds one frame group only
    ex = olink.code.export(with_enc=False)
    Z, layers = minsum_ref.qc_structure(ex["row_ptr"], ex["col_idx"], len(ex["row_ptr"]) - 1, olink.code.N)
    assert Z == 96 and len(layers) == 12
    llr = np.stack([util.llr_of_p0(r.p0) for r in rs]).astype(np.float32)
    link = util.gpu_link(name, algorithm=3)
    for alpha, beta in ((0.8, 0.0), (1.0, 0.5)):
        link.set_minsum(alpha, beta)
        cc, uu, ret = link.decode(llr)
        rret, rcc = minsum_ref.decode_layered(layers, Z, olink.code.N, olink.code.two_z, llr, 50, 50, alpha=alpha, beta=beta)
        assert np.array_equal(ret, rret), np.where(ret != rret)[0][:8]
        assert np.array_equal(cc, rcc)
    # fixed-iteration mode latches the same answers
    link.set_early_exit(False)
    cc_fix, _, ret_fix = link.decode(llr)
    assert np.array_equal(ret_fix, ret) and np.array_equal(cc_fix[ret < 50], cc[ret < 50])
    link.close()


@pytest.mark.parametrize("snr,frames", [(10.0, 20000), (12.0, 20000)])
def test_layered_minsum_gate(snr, frames):
    """BER/FER of the layered decoder against the sum-product decoder on the same Philox frames.  (Iterations: in this
    link most decoder time goes to frames the blind detector got wrong — they run all 50 iterations under any schedule —
    so the totals barely move; the schedule's advantage is asserted on converging frames below.)"""
    link = util.gpu_link("5g_16qam_gray_10db", max_batch=8192)
    spa, it_spa = link.simulate(snr, frames, seed=37)
    link.set_algorithm(1, 0.8)
    ms, it_ms = link.simulate(snr, frames, seed=37)
    link.set_algorithm(3, 0.8)
    lay, it_lay = link.simulate(snr, frames, seed=37)
    fer_spa, fer_lay = spa[1] / spa[0], lay[1] / lay[0]
    sd = np.sqrt(max(fer_spa * (1 - fer_spa), 1e-4) / frames)
    assert fer_spa - 3 * sd - 0.005 <= fer_lay <= fer_spa + 3 * sd + 0.01, (fer_lay, fer_spa)
    assert abs(lay[3] / lay[2] - spa[3] / spa[2]) <= 0.1 * spa[3] / spa[2] + 2e-3
    assert it_lay <= it_ms, (it_lay, it_ms, it_spa)
    # converging frames: fewer iterations than the flooding min-sum, although the layered stop rule needs one clean
    # iteration after the decisions became a codeword
    u, c, h, y = link.generate(2000, snr, seed=39)
    llr = link.demap(y, h, 10 ** (-0.1 * snr))
    _, _, ret_lay = link.decode(llr)
    link.set_algorithm(1, 0.8)
    _, _, ret_ms = link.decode(llr)
    both = (ret_lay < 50) & (ret_ms < 50) & (ret_ms > 2)
    assert both.sum() > 200 and ret_lay[both].sum() < 0.9 * ret_ms[both].sum(), (ret_lay[both].sum(), ret_ms[both].sum())
    link.set_algorithm(0)
    again, _ = link.simulate(snr, frames, seed=37)
    assert np.array_equal(again, spa)
    link.close()


def test_layered_needs_a_quasi_cyclic_code():
    import kmldpc_b200 as kb
    link = util.gpu_link("peg2304_4psk_6db")
    with pytest.raises(kb.KmlError, match="layered"):
        link.set_algorithm(3)
    link.close()
    with pytest.raises(kb.KmlError, match="layered"):
        util.gpu_link("peg2304_4psk_6db", algorithm=3)


def test_sweep_driver_accepts_the_throughput_algorithms(tmp_path):
    """`[gpu] algorithm` in config.toml reaches the decoder through kml_sweep_run: the 5G sweep with the layered and the
    flooding min-sum gives the sum-product's FER within the gate; algorithm 3 on a code without quasi-cyclic structure is an
    error with a message, not a crash."""
    import kmldpc_b200 as kb

    def run(matrix, is5g, alg):
        cfg = tmp_path / f"cfg_{alg}_{int(is5g)}.toml"
        cfg.write_text(f"""[range]
minimum_snr = 10.0
maximum_snr = 12.0
step_snr = 2.0
maximum_error_number = 100000000
maximum_block_number = 8192
thread_block_number = 32
[decoder]
true_h_arg = false
[xcodec]
5gldpc = {"true" if is5g else "false"}
metric_type = false
metric_iter = 5
[histogram]
enable = false
[ldpc]
max_iter = 50
active = true
matrix_file = "{matrix}"
[modem]
modem_file = "4bit_16QAM_Gray.txt"
[gpu]
seed = 17
batch = 2048
algorithm = {alg}
""")
        return kb.Simulator(str(cfg), data_dir=util.ko.CONFIG_DIR).simulate(echo=False)

    _, _, fer0, cnt0 = run("5GLDPCBG2a3_R12_K960.txt", True, 0)
    for alg in (1, 3):
        _, _, fer, cnt = run("5GLDPCBG2a3_R12_K960.txt", True, alg)
        assert (cnt[:, 0] == 8192).all()
        assert np.all(np.abs(fer - fer0) <= 0.03), (alg, fer, fer0)
    with pytest.raises(kb.KmlError, match="layered"):
        run("PEG2304regular0.5.txt", False, 3)
