"""CPU-side tests (no GPU): the shared library loads and exports every symbol include/kmldpc_b200.h declares, the
host-side code/constellation construction equals the oracle (hence the reference), config parsing, and the product
path refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import kmldpc_b200 as kb
from kmldpc_b200 import capi
from oracle import kml_oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "kmldpc_b200.h")).read()
    declared = set(re.findall(r"\b(kml_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"kml_status"}
    lib = capi.load()
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in capi.SYMBOLS, f"{name} has no ctypes prototype"
    assert set(capi.SYMBOLS) <= declared


@pytest.mark.parametrize("matrix,is5g,active", [("PEG2304regular0.5.txt", False, True),
                                                ("PEG8064regular0.5.txt", False, True),
                                                ("5GLDPCBG2a3_R12_K960.txt", True, True),
                                                ("PEG2304regular0.5.txt", False, False)])
def test_code_construction_equals_oracle(matrix, is5g, active):
    code = kb.LdpcCode(matrix, is_5g=is5g, active=active)
    oc = ko.Code(matrix, is5g, active)
    ex = oc.export()
    assert (code.M, code.N, code.N_tx, code.K, code.n_chk, code.puncture) == (oc.M, oc.N, oc.N_tx, oc.K, oc.chk, oc.two_z)
    assert code.info_offset == (0 if is5g else oc.chk)
    assert np.array_equal(code.perm, ex["perm"])
    rp, ci = ex["row_ptr"], ex["col_idx"]
    for r in range(code.M):
        assert sorted(ci[rp[r]:rp[r + 1]]) == list(code.col_idx[code.row_ptr[r]:code.row_ptr[r + 1]])
    if active:
        enc = ex["enc_h"]
        info = enc[:, :code.K] if is5g else enc[:, code.n_chk:]
        assert np.array_equal(kb.unpack_bits(code.enc_rows, code.K), info)
        ident = enc[:, code.K:] if is5g else enc[:, :code.n_chk]
        assert np.array_equal(ident, np.eye(code.M, dtype=np.uint8))


@pytest.mark.parametrize("f", ["2bits_QPSK.txt", "2bits_4PSK.txt", "4bit_16QAM_Gray.txt", "4bit_16QAM_phi1.txt",
                               "4bit_16QAM_phi2.txt", "6bits_64QAM_Gray.txt"])
def test_constellations_equal_oracle(f):
    m, om = kb.Modem(f), ko.Modem(f)
    assert (m.bits, m.Q) == (om.bits, om.Q)
    assert np.abs(m.points - om.points).max() <= 1e-15
    assert abs((np.abs(m.points) ** 2).mean() - 1.0) < 1e-12


def test_missing_files_return_errors_not_exit():
    with pytest.raises(kb.KmlError, match="cannot open"):
        kb.LdpcCode("/nonexistent/H.txt")
    with pytest.raises(kb.KmlError, match="cannot open"):
        kb.Modem("/nonexistent/c.txt")


def test_config_toml_parsing_matches_reference_keys(tmp_path):
    sim = kb.Simulator(os.path.join(ROOT, "config", "config.toml"))
    c = sim.cfg
    assert (c.min_snr, c.max_snr, c.step_snr) == (15.0, 15.0, 5.0)
    assert (c.max_err_blk, c.max_num_blk, c.known_h, c.is_5g, c.metric_type, c.metric_iter) == (1, 1, 0, 0, 0, 5)
    assert (c.max_iter, c.encoder_active, c.histogram_enable) == (50, 1, 0)
    assert c.matrix_file == b"PEG2304regular0.5.txt" and c.modem_file == b"4bit_16QAM_Gray.txt"
    assert sim.n_points == 1 and c.seed == 17 and c.n_gpus == 1
    bad = tmp_path / "bad.toml"
    bad.write_text("[range]\nminimum_snr = 1.0\n")
    with pytest.raises(kb.KmlError, match="missing key"):
        kb.Simulator(str(bad))
    # (unsigned long)((max - min) / step + 1), simulator.cc:27
    sim2 = kb.Simulator(os.path.join(ROOT, "config", "config.toml"), min_snr=0.0, max_snr=30.0, step_snr=5.0)
    assert sim2.n_points == 7
    # the optional [gpu] table (ignored by the reference binary): defaults, then every key
    assert (c.reduce_on_host, c.debug_frames, c.early_exit, c.algorithm, c.max_batch) == (0, 0, 1, 0, 0)
    full = tmp_path / "gpu.toml"
    full.write_text(open(os.path.join(ROOT, "config", "config.toml")).read() +
                    '\n[gpu]\nseed = 99\ngpus = 4\nbatch = 4096\nearly_exit = false\nalgorithm = 1\nreduce = "host"\ndebug = true\n')
    g = kb.Simulator(str(full)).cfg
    assert (g.seed, g.n_gpus, g.max_batch, g.early_exit, g.algorithm, g.reduce_on_host, g.debug_frames) == (99, 4, 4096, 0, 1, 1, 1)


@pytest.mark.gpu
def test_sweep_with_zero_error_budget_runs_nothing_like_the_reference(tmp_path):
    """maximum_error_number = 0: simulator.cc:117 leaves before the first frame (err_blk >= 0 holds at once)."""
    cfg = tmp_path / "z.toml"
    cfg.write_text(open(os.path.join(ROOT, "config", "config.toml")).read().replace("maximum_error_number = 1", "maximum_error_number = 0"))
    snr, ber, fer, cnt = kb.Simulator(str(cfg), data_dir=os.path.join(ROOT, "config")).simulate(echo=False)
    assert cnt[0, 0] == 0 and np.isnan(ber[0]) and np.isnan(fer[0])


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kb.KmlError, match="no CUDA device|CUDA"):
        kb.Link(kb.LdpcCode("PEG2304regular0.5.txt"), kb.Modem("2bits_QPSK.txt"))


def test_bit_packing_helpers_roundtrip():
    rng = np.random.default_rng(0)
    b = rng.integers(0, 2, size=(5, 1152), dtype=np.uint8)
    p = kb.pack_bits(b)
    assert p.shape == (5, 36) and np.array_equal(kb.unpack_bits(p, 1152), b)
    assert p[0, 0] & 1 == b[0, 0] and (p[0, 1] >> 3) & 1 == b[0, 35]


@pytest.mark.parametrize("m,n,col_degs,seed,dup", [(150, 334, (1, 2, 3, 4), 1, False), (120, 300, (2, 3, 3, 3, 12), 2, False),
                                                   (64, 200, (2, 3), 7, True)])
def test_code_construction_on_random_irregular_matrices(tmp_path, m, n, col_degs, seed, dup):
    """The bit-packed elimination against the oracle's byte-matrix restatement away from the three shipped matrices:
    irregular degrees, an all-zero column, N not a multiple of 32, and (dup) a rank-deficient H with a repeated row."""
    from tests.test_gpu_parity import _write_irregular_code
    path = str(tmp_path / "irr.txt")
    _write_irregular_code(path, m, n, col_degs, seed)
    if dup:  # repeat row 3 as a new last row: rank stays m, the file declares m + 1 rows
        lines = open(path).read().splitlines()
        row3 = lines[3 + 3].split()
        lines[1] = "%d\t%d\t%d" % (m + 1, n, m)
        lines.append(" ".join([str(m)] + row3[1:]) + " ")
        open(path, "w").write("\n".join(lines) + "\n")
    code, oc = kb.LdpcCode(path), ko.Code(path, False, True)
    ex = oc.export()
    assert (code.M, code.N, code.K, code.n_chk) == (oc.M, oc.N, oc.K, oc.chk)
    assert np.array_equal(code.perm, ex["perm"])
    rp, ci = ex["row_ptr"], ex["col_idx"]
    for r in range(code.M):
        assert sorted(ci[rp[r]:rp[r + 1]]) == list(code.col_idx[code.row_ptr[r]:code.row_ptr[r + 1]])
    assert np.array_equal(kb.unpack_bits(code.enc_rows, code.K), ex["enc_h"][:code.n_chk, code.n_chk:])
