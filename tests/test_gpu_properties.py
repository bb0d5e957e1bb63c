"""Size-independent properties at BASELINE.json's full sizes (batches of 16384 frames, 10^5-frame BER points), where the
CPU oracle would take minutes: encoder linearity and zero syndrome, noiseless round trips through the whole link,
decoder idempotence, batch/split invariance, and the reference's high-SNR behaviour per constellation."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kb():
    import kmldpc_b200
    return kmldpc_b200


@pytest.mark.parametrize("name", ["peg2304_qpsk_10db", "peg8064_64qam_20db", "5g_16qam_gray_10db"])
def test_encoder_linearity_and_zero_syndrome_full_batch(name, kb):
    B = 4096 if "8064" in name else 16384
    link = util.gpu_link(name, max_batch=B)
    K, N, punct = link.code.K, link.code.N, link.code.puncture
    rng = np.random.default_rng(7)
    u1 = rng.integers(0, 2, size=(B, K), dtype=np.int32)
    u2 = rng.integers(0, 2, size=(B, K), dtype=np.int32)
    c1, c2, c12 = link.encode(u1), link.encode(u2), link.encode(u1 ^ u2)
    assert np.array_equal(c1 ^ c2, c12)                                   # GF(2) linearity
    assert not link.encode(np.zeros((3, K), np.int32)).any()
    # systematic part: PEG → info on the right of the word, 5G → info on the left, first 2Z bits punctured
    if link.code.is_5g:
        assert np.array_equal(c1[:, :K - punct], u1[:, punct:])
    else:
        assert np.array_equal(c1[:, link.code.info_offset:], u1)
    # every codeword has zero syndrome: a confident LLR input leaves the decoder at t = 0 with the same word
    if not link.code.is_5g:
        llr = (1.0 - 2.0 * c1[:2048]).astype(np.float32) * 20.0
        cc, uu, ret = link.decode(llr)
        assert np.array_equal(cc, c1[:2048]) and (ret == 1).all() and np.array_equal(uu, u1[:2048])
    link.close()


@pytest.mark.parametrize("name,bits", [("peg2304_4psk_6db", 2), ("peg2304_16qam_gray_12db", 4),
                                       ("peg8064_64qam_20db", 6), ("5g_16qam_gray_10db", 4)])
def test_noiseless_link_round_trip(name, bits, kb):
    """generate (80 dB) → blind receive → decoded bits equal the source bits on every frame, at iteration 0."""
    B = 2048
    link = util.gpu_link(name, max_batch=B)
    u, c, h, y = link.generate(B, 80.0, seed=3)
    uu_p, hhat, kstar, ret = link.receive(y, 1e-8)
    uu = kb.unpack_bits(uu_p, link.code.K)
    # blind detection resolves h only up to the rotation the syndrome picks; Gray/4PSK labelings decode correctly
    ok = (uu == u).all(axis=1)
    assert ok.mean() == 1.0, float(ok.mean())
    assert (ret == 1).all() or link.code.is_5g
    rot = np.exp(1j * (3.14159265358979 / 2) * kstar)
    # the estimate is the mean of the cluster-0 samples: its error is noise / sqrt(count), tiny at 80 dB except for deep fades
    assert np.median(np.abs(hhat * rot - h) / np.abs(h)) < 1e-4 and (np.abs(hhat * rot - h) / np.abs(h)).max() < 2e-2
    link.close()


def test_qpsk_file_floor_and_phi_mappings_at_high_snr(kb):
    """The reference's quirks at full scale: the 2bits_QPSK labeling cannot tell 0° from 180° (all-ones is a
    codeword) → FER ≈ 0.5 at any SNR; phi1 blind → FER = 1; known-h phi1 decodes."""
    B = 8192
    link = util.gpu_link("peg2304_qpsk_10db", max_batch=B)
    cnt, _ = link.simulate(40.0, B, seed=11)
    assert 0.45 < cnt[1] / cnt[0] < 0.55 and 0.45 < cnt[3] / cnt[2] < 0.55
    link.close()
    link = util.gpu_link("peg2304_16qam_phi1_15db", max_batch=B)
    cnt, _ = link.simulate(40.0, 2048, seed=11)
    assert cnt[1] == cnt[0]
    link.close()
    code, mod = kb.LdpcCode("PEG2304regular0.5.txt"), kb.Modem("4bit_16QAM_phi1.txt")
    link = kb.Link(code, mod, known_h=True, max_batch=B)
    cnt, _ = link.simulate(40.0, 2048, seed=11)
    assert cnt[1] / cnt[0] < 0.02
    link.close()


def test_decoder_idempotence_and_determinism_full_batch(kb):
    B = 16384
    link = util.gpu_link("peg2304_4psk_6db", max_batch=B)
    u, c, h, y = link.generate(B, 6.0, seed=21)
    llr = link.demap(y, h, 10 ** -0.6)                      # true h
    cc, uu, ret = link.decode(llr)
    cc2, uu2, ret2 = link.decode(llr)
    assert np.array_equal(cc, cc2) and np.array_equal(ret, ret2)          # dynamic frame queue, same answers
    conv = ret < 50
    assert 0.3 < conv.mean() < 1.0
    # a converged word fed back as a confident LLR is a fixed point
    back = (1.0 - 2.0 * cc[conv][:4096]).astype(np.float32) * 15.0
    cc3, _, ret3 = link.decode(back)
    assert np.array_equal(cc3, cc[conv][:4096]) and (ret3 == 1).all()
    # decoded = transmitted on the frames without bit errors; error frames never report a codeword equal to c
    same = (cc == c).all(axis=1)
    assert np.array_equal(same, (uu == u).all(axis=1) & same)
    # fixed-iteration mode: identical decisions and return values on all 16384 frames
    link.set_early_exit(False)
    cc4, _, ret4 = link.decode(llr)
    assert np.array_equal(ret4, ret) and np.array_equal(cc4[conv], cc[conv])
    link.close()


def test_ber_point_1e5_frames_split_invariance(kb):
    """10^5 frames of C1 at one SNR point: counters are the same whether run in one call or as 7 ragged pieces
    (Philox counters come from the global frame index), and FER/BER decrease with SNR."""
    link = util.gpu_link("peg2304_4psk_6db", max_batch=8192)
    total = 100000
    whole, it = link.simulate(8.0, total, seed=17)
    parts = np.zeros(4, np.uint64)
    edges = [0, 1, 4097, 33333, 33334, 70001, 99999, total]
    for a, b in zip(edges, edges[1:]):
        c, _ = link.simulate(8.0, b - a, seed=17, frame_begin=a)
        parts += c
    assert np.array_equal(whole, parts) and whole[0] == total
    lo, _ = link.simulate(4.0, 20000, seed=17)
    hi, _ = link.simulate(14.0, 20000, seed=17)
    assert lo[1] / lo[0] > whole[1] / whole[0] > hi[1] / hi[0]
    assert lo[3] / lo[2] > whole[3] / whole[2] > hi[3] / hi[2]
    link.close()
