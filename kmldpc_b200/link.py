"""Host-side mirror of the reference's link objects on top of the C ABI.

  LdpcCode      ~ lab::BinaryLDPCCodec / lab::Binary5GLDPCCodec construction (H file → permuted graph + encoder)
  Modem         ~ lab::Modem::init
  Link          ~ KmCodec + ModemLinearSystem + KMeans as used by Simulator::run_blocks (src/simulator.cc:112-168):
                  encode / generate / modulate / kmeans / demap / resolve / decode / receive / simulate

Arrays are numpy host arrays (the C ABI takes host pointers); `*_dev` methods take raw device pointers (ints) for
callers that keep data in HBM (bench.py).  Everything computes on the GPU — there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import capi

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config")


class KmlError(RuntimeError):
    pass


def _resolve(path: str) -> str:
    return path if os.path.isabs(path) or os.path.exists(path) else os.path.join(CONFIG_DIR, path)


def _ptr(a: np.ndarray | None, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


class LdpcCode:
    """Parity-check file → code description (host only; runs the bit-packed Gaussian elimination)."""

    def __init__(self, h_file: str, is_5g: bool = False, active: bool = True):
        self._lib = capi.load()
        self._p = C.POINTER(capi.KmlCode)()
        rc = self._lib.kml_code_load(_resolve(h_file).encode(), int(is_5g), int(active), C.byref(self._p))
        if rc != 0:
            raise KmlError(f"kml_code_load({h_file}): {self._lib.kml_last_error(None).decode()}")
        d = self._p.contents
        self.M, self.N, self.N_tx, self.K = d.n_rows, d.n_graph, d.n_tx, d.k
        self.n_chk, self.puncture, self.info_offset, self.E = d.n_chk, d.puncture, d.info_offset, d.n_edges
        self.is_5g, self.active, self.enc_words = bool(d.is_5g), bool(d.encoder_active), d.enc_words

    @property
    def row_ptr(self):
        return np.ctypeslib.as_array(self._p.contents.row_ptr, (self.M + 1,)).copy()

    @property
    def col_idx(self):
        return np.ctypeslib.as_array(self._p.contents.col_idx, (self.E,)).copy()

    @property
    def perm(self):
        return np.ctypeslib.as_array(self._p.contents.perm, (self.N,)).copy()

    @property
    def enc_rows(self):
        if not self.active:
            return None
        return np.ctypeslib.as_array(self._p.contents.enc_rows, (self.n_chk, self.enc_words)).copy()

    def __del__(self):
        if getattr(self, "_p", None):
            self._lib.kml_code_free(self._p)
            self._p = None


class Modem:
    def __init__(self, modem_file: str):
        self._lib = capi.load()
        self._p = C.POINTER(capi.KmlModem)()
        rc = self._lib.kml_modem_load(_resolve(modem_file).encode(), C.byref(self._p))
        if rc != 0:
            raise KmlError(f"kml_modem_load({modem_file}): {self._lib.kml_last_error(None).decode()}")
        d = self._p.contents
        self.bits, self.Q = d.bits_per_symbol, d.n_points
        self.points = np.ctypeslib.as_array(d.points, (self.Q, 2)).copy().view(np.complex128).reshape(self.Q)

    def __del__(self):
        if getattr(self, "_p", None):
            self._lib.kml_modem_free(self._p)
            self._p = None


class Link:
    def __init__(self, code: LdpcCode, modem: Modem, *, max_iter=50, known_h=False, metric_type=False, metric_iter=5,
                 kmeans_iter=20, early_exit=True, max_batch=0, device=0, algorithm=0):
        self._lib = capi.load()
        self.code, self.modem = code, modem
        self.opts = capi.KmlOpts(max_iter, int(known_h), int(metric_type), metric_iter, kmeans_iter, int(early_exit),
                                 max_batch, int(algorithm))
        self._h = C.c_void_p()
        rc = self._lib.kml_create(C.byref(self._h), device, code._p, modem._p, C.byref(self.opts))
        if rc != 0:
            raise KmlError(f"kml_create: {self._lib.kml_last_error(None).decode()} (rc={rc})")
        info = (C.c_int32 * 8)()
        self._lib.kml_info(self._h, info)
        self.n_sym, self.max_batch = info[6], info[7]
        self.k_words = (code.K + 31) // 32
        self.words_n = (code.N + 31) // 32
        self.max_iter, self.known_h = max_iter, bool(known_h)

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self._lib.kml_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise KmlError(f"{what}: {self._lib.kml_last_error(self._h).decode()} (rc={rc})")

    def decoder_info(self) -> dict:
        info = (C.c_int32 * 8)()
        self._check(self._lib.kml_decoder_info(self._h, info), "kml_decoder_info")
        keys = ("kernel_kind", "threads", "smem_bytes", "ctas_per_sm", "layout_residual", "excess_wavefronts",
                "gather_instructions", "row_slots")
        d = dict(zip(keys, list(info)))
        d["qc_plan"], d["row_major"] = (d["kernel_kind"] >> 8) & 0xFF, (d["kernel_kind"] >> 16) & 1
        d["kernel_kind"] &= 0xFF
        return d

    @property
    def launches(self) -> int:
        return int(self._lib.kml_launch_count(self._h))

    def set_early_exit(self, flag: bool):
        self._check(self._lib.kml_set_early_exit(self._h, int(flag)), "kml_set_early_exit")

    def set_algorithm(self, algorithm: int, alpha: float = 0.8):
        """0 = sum-product (the reference's decoder), 1 = normalised min-sum (throughput mode, not reference-pinned)."""
        self._check(self._lib.kml_set_algorithm(self._h, int(algorithm), float(alpha)), "kml_set_algorithm")

    def set_minsum(self, alpha: float = 0.8, beta: float = 0.0):
        """Check-node rule of the min-sum decoders: |c2v| = max(alpha * min - beta, 0) (normalised / offset / both)."""
        self._check(self._lib.kml_set_minsum(self._h, float(alpha), float(beta)), "kml_set_minsum")

    # ---- stages (host buffers)
    def encode(self, u: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(u, np.int32).reshape(-1, self.code.K)
        c = np.empty((u.shape[0], self.code.N_tx), np.int32)
        self._check(self._lib.kml_encode(self._h, u.shape[0], _ptr(u, C.c_int32), _ptr(c, C.c_int32)), "kml_encode")
        return c

    def generate(self, B: int, snr_db: float, seed: int = 17, frame0: int = 0):
        u = np.empty((B, self.code.K), np.int32)
        c = np.empty((B, self.code.N_tx), np.int32)
        h = np.empty((B, 2), np.float32)
        y = np.empty((B, self.n_sym, 2), np.float32)
        self._check(self._lib.kml_generate(self._h, B, snr_db, seed, frame0, _ptr(u, C.c_int32), _ptr(c, C.c_int32),
                                           _ptr(h, C.c_float), _ptr(y, C.c_float)), "kml_generate")
        return u, c, h.view(np.complex64).reshape(B), y.view(np.complex64).reshape(B, self.n_sym)

    def modulate(self, c: np.ndarray, h: np.ndarray, noise: np.ndarray | None, sigma: float) -> np.ndarray:
        c = np.ascontiguousarray(c, np.int32).reshape(-1, self.code.N_tx)
        B = c.shape[0]
        h = np.ascontiguousarray(np.asarray(h, np.complex64).reshape(B)).view(np.float32)
        nz = None if noise is None else np.ascontiguousarray(np.asarray(noise, np.complex64).reshape(B, self.n_sym)).view(np.float32)
        y = np.empty((B, self.n_sym, 2), np.float32)
        self._check(self._lib.kml_modulate(self._h, B, _ptr(c, C.c_int32), _ptr(h, C.c_float), _ptr(nz, C.c_float),
                                           sigma, _ptr(y, C.c_float)), "kml_modulate")
        return y.view(np.complex64).reshape(B, self.n_sym)

    def _y(self, y):
        y = np.ascontiguousarray(np.asarray(y, np.complex64).reshape(-1, self.n_sym))
        return y, y.view(np.float32)

    def kmeans(self, y: np.ndarray):
        y, yf = self._y(y)
        B = y.shape[0]
        hhat = np.empty((B, 2), np.float32)
        passes = np.empty(B, np.int32)
        self._check(self._lib.kml_kmeans(self._h, B, _ptr(yf, C.c_float), _ptr(hhat, C.c_float), _ptr(passes, C.c_int32)),
                    "kml_kmeans")
        return hhat.view(np.complex64).reshape(B), passes

    def kmeans_f64(self, y: np.ndarray):
        """The reference's types: y complex128 [B, n_sym] → hhat complex128 [B] (the estimate as carried in fp64)."""
        y = np.ascontiguousarray(np.asarray(y, np.complex128).reshape(-1, self.n_sym))
        B = y.shape[0]
        hhat = np.empty((B, 2), np.float64)
        passes = np.empty(B, np.int32)
        self._check(self._lib.kml_kmeans_f64(self._h, B, _ptr(y.view(np.float64), C.c_double), _ptr(hhat, C.c_double),
                                             _ptr(passes, C.c_int32)), "kml_kmeans_f64")
        return hhat.view(np.complex128).reshape(B), passes

    def demap(self, y: np.ndarray, h: np.ndarray, var: float) -> np.ndarray:
        y, yf = self._y(y)
        B = y.shape[0]
        hf = np.ascontiguousarray(np.asarray(h, np.complex64).reshape(B)).view(np.float32)
        llr = np.empty((B, self.code.N_tx), np.float32)
        self._check(self._lib.kml_demap(self._h, B, _ptr(yf, C.c_float), _ptr(hf, C.c_float), var, _ptr(llr, C.c_float)),
                    "kml_demap")
        return llr

    def resolve(self, y: np.ndarray, hhat: np.ndarray, var: float):
        y, yf = self._y(y)
        B = y.shape[0]
        hf = np.ascontiguousarray(np.asarray(hhat, np.complex64).reshape(B)).view(np.float32)
        metric = np.empty((B, 4), np.float32)
        kstar = np.empty(B, np.int32)
        self._check(self._lib.kml_resolve(self._h, B, _ptr(yf, C.c_float), _ptr(hf, C.c_float), var,
                                          _ptr(metric, C.c_float), _ptr(kstar, C.c_int32)), "kml_resolve")
        return metric, kstar

    def decode(self, llr: np.ndarray, iter_count: int | None = None):
        llr = np.ascontiguousarray(llr, np.float32).reshape(-1, self.code.N_tx)
        B = llr.shape[0]
        cc = np.empty((B, self.code.N), np.int32)
        uu = np.empty((B, self.code.K), np.int32)
        ret = np.empty(B, np.int32)
        self._check(self._lib.kml_decode(self._h, B, _ptr(llr, C.c_float), iter_count or self.max_iter,
                                         _ptr(cc, C.c_int32), _ptr(uu, C.c_int32), _ptr(ret, C.c_int32)), "kml_decode")
        return cc, uu, ret

    def decode_p0(self, p0: np.ndarray, iter_count: int | None = None):
        """BinaryLDPCCodec::Decoder's own input: P(bit = 0) in double [B, n_tx]."""
        p0 = np.ascontiguousarray(p0, np.float64).reshape(-1, self.code.N_tx)
        B = p0.shape[0]
        cc = np.empty((B, self.code.N), np.int32)
        uu = np.empty((B, self.code.K), np.int32)
        ret = np.empty(B, np.int32)
        self._check(self._lib.kml_decode_p0(self._h, B, _ptr(p0, C.c_double), iter_count or self.max_iter,
                                            _ptr(cc, C.c_int32), _ptr(uu, C.c_int32), _ptr(ret, C.c_int32)), "kml_decode_p0")
        return cc, uu, ret

    def receive_f64(self, y: np.ndarray, var: float, true_h: np.ndarray | None = None, with_metric: bool = False):
        """kml_receive on the reference's types: y complex128 [B, n_sym], true_h complex128 [B].
        Returns uu_hat_packed, hhat (complex128), kstar, ret[, metric[B, 4]]."""
        y = np.ascontiguousarray(np.asarray(y, np.complex128).reshape(-1, self.n_sym))
        B = y.shape[0]
        th = None
        if true_h is not None:
            th = np.ascontiguousarray(np.asarray(true_h, np.complex128).reshape(B)).view(np.float64)
        uu = np.empty((B, self.k_words), np.uint32)
        hhat = np.zeros((B, 2), np.float64)
        kstar = np.zeros(B, np.int32)
        ret = np.empty(B, np.int32)
        met = np.zeros((B, 4), np.float32) if with_metric else None
        self._check(self._lib.kml_receive_f64(self._h, B, _ptr(y.view(np.float64), C.c_double), _ptr(th, C.c_double), var,
                                              _ptr(uu, C.c_uint32), _ptr(hhat, C.c_double), _ptr(kstar, C.c_int32),
                                              _ptr(ret, C.c_int32), _ptr(met, C.c_float)), "kml_receive_f64")
        out = (uu, hhat.view(np.complex128).reshape(B), kstar, ret)
        return out + (met,) if with_metric else out

    @property
    def soft_state(self) -> float:
        """Sum of ln(syndrom_soft_) as the context's last Decoder call left it (soft-syndrome metric's stale-value chain)."""
        v = C.c_double(0.0)
        self._check(self._lib.kml_soft_syndrome_state(self._h, 0, C.byref(v)), "kml_soft_syndrome_state")
        return float(v.value)

    @soft_state.setter
    def soft_state(self, value: float):
        v = C.c_double(float(value))
        self._check(self._lib.kml_soft_syndrome_state(self._h, 1, C.byref(v)), "kml_soft_syndrome_state")

    def receive(self, y: np.ndarray, var: float, true_h: np.ndarray | None = None, out=None, with_metric: bool = False):
        """y: complex64 [B, n_sym] (or a float32 view).  Returns uu_hat_packed, hhat, kstar, ret[, metric[B, 4]]."""
        y, yf = self._y(y)
        B = y.shape[0]
        th = None
        if true_h is not None:
            th = np.ascontiguousarray(np.asarray(true_h, np.complex64).reshape(B)).view(np.float32)
        uu = np.empty((B, self.k_words), np.uint32) if out is None else out
        hhat = np.zeros((B, 2), np.float32)
        kstar = np.zeros(B, np.int32)
        ret = np.empty(B, np.int32)
        met = np.zeros((B, 4), np.float32) if with_metric else None
        self._check(self._lib.kml_receive(self._h, B, _ptr(yf, C.c_float), _ptr(th, C.c_float), var,
                                          _ptr(uu, C.c_uint32), _ptr(hhat, C.c_float), _ptr(kstar, C.c_int32),
                                          _ptr(ret, C.c_int32), _ptr(met, C.c_float)), "kml_receive")
        res = (uu, hhat.view(np.complex64).reshape(B), kstar, ret)
        return res + (met,) if with_metric else res

    def receive_raw(self, B: int, y_ptr: int, var: float, uu_ptr: int, ret_ptr: int = 0, true_h_ptr: int = 0):
        """Host pointers given as integers (e.g. pinned torch tensors): no numpy wrapping, no allocation."""
        f = self._lib.kml_receive
        self._check(f(self._h, B, C.cast(y_ptr, capi.c_f32p), C.cast(true_h_ptr, capi.c_f32p) if true_h_ptr else None, var,
                      C.cast(uu_ptr, capi.c_u32p), None, None, C.cast(ret_ptr, capi.c_i32p) if ret_ptr else None, None),
                    "kml_receive")

    def receive_submit_raw(self, B: int, y_ptr: int, var: float, uu_ptr: int, ret_ptr: int = 0):
        """kml_receive_submit on host pointers given as integers: enqueue the batch and return (see receive_wait)."""
        self._check(self._lib.kml_receive_submit(self._h, B, C.cast(y_ptr, capi.c_f32p), None, var, C.cast(uu_ptr, capi.c_u32p), None,
                                                 None, C.cast(ret_ptr, capi.c_i32p) if ret_ptr else None, None), "kml_receive_submit")

    def receive_wait(self, max_outstanding: int = 0):
        """Block until at most `max_outstanding` submitted batches are still in flight."""
        self._check(self._lib.kml_receive_wait(self._h, int(max_outstanding)), "kml_receive_wait")

    def receive_f64_raw(self, B: int, y_ptr: int, var: float, uu_ptr: int, ret_ptr: int = 0):
        """kml_receive_f64 on host pointers given as integers (pinned complex128 symbols): no wrapping, no allocation."""
        self._check(self._lib.kml_receive_f64(self._h, B, C.cast(y_ptr, capi.c_f64p), None, var, C.cast(uu_ptr, capi.c_u32p), None,
                                              None, C.cast(ret_ptr, capi.c_i32p) if ret_ptr else None, None), "kml_receive_f64")

    def count_errors(self, u_packed: np.ndarray, uu_hat_packed: np.ndarray):
        u = np.ascontiguousarray(u_packed, np.uint32).reshape(-1, self.k_words)
        uh = np.ascontiguousarray(uu_hat_packed, np.uint32).reshape(-1, self.k_words)
        cnt = np.zeros(4, np.uint64)
        self._check(self._lib.kml_count_errors(self._h, u.shape[0], _ptr(u, C.c_uint32), _ptr(uh, C.c_uint32),
                                               _ptr(cnt, C.c_uint64)), "kml_count_errors")
        return cnt

    def simulate(self, snr_db: float, frames: int, *, seed: int = 17, frame_begin: int = 0, max_err_blk: int = 0):
        """Returns (counters[4] = tot_blk, err_blk, tot_bit, err_bit, iterations executed)."""
        cnt = np.zeros(4, np.uint64)
        it = C.c_uint64(0)
        self._check(self._lib.kml_simulate(self._h, snr_db, seed, frame_begin, frames, max_err_blk,
                                           _ptr(cnt, C.c_uint64), C.byref(it)), "kml_simulate")
        return cnt, int(it.value)

    def measure_smem_bandwidth(self) -> float:
        """GB/s of conflict-free LDS.128 over all SMs of this device (the decoder's roofline denominator)."""
        v = C.c_double(0.0)
        self._check(self._lib.kml_measure_smem_bandwidth(self._h, C.byref(v)), "kml_measure_smem_bandwidth")
        return float(v.value)

    def histogram(self, snr_db: float, frames: int, *, seed: int = 17, frame_begin: int = 0):
        """Histogram mode (simulator.cc:154-162): the four candidate metrics per frame + the counters the reference
        accumulates in that mode."""
        met = np.empty((frames, 4), np.float32)
        cnt = np.zeros(4, np.uint64)
        self._check(self._lib.kml_histogram(self._h, snr_db, seed, frame_begin, frames, _ptr(met, C.c_float),
                                            _ptr(cnt, C.c_uint64)), "kml_histogram")
        return met, cnt

    def histogram_rx(self, y: np.ndarray, var: float, u_packed: np.ndarray):
        """KmCodec::GetHistogramData + CntErr on given frames: metrics[B,4], kstar, uu_hat_packed (what CntErr saw), counters."""
        y, yf = self._y(y)
        B = y.shape[0]
        u = np.ascontiguousarray(u_packed, np.uint32).reshape(B, self.k_words)
        met = np.empty((B, 4), np.float32)
        kstar = np.empty(B, np.int32)
        uh = np.empty((B, self.k_words), np.uint32)
        cnt = np.zeros(4, np.uint64)
        self._check(self._lib.kml_histogram_rx(self._h, B, _ptr(yf, C.c_float), var, _ptr(u, C.c_uint32), _ptr(met, C.c_float),
                                               _ptr(kstar, C.c_int32), _ptr(uh, C.c_uint32), _ptr(cnt, C.c_uint64)),
                    "kml_histogram_rx")
        return met, kstar, uh, cnt

    # ---- device-pointer variants (pointers as ints, stream as int)
    def generate_dev(self, B, snr_db, seed, frame0, u_packed_ptr, h_ptr, y_ptr, stream=0):
        self._check(self._lib.kml_generate_dev(self._h, B, snr_db, seed, frame0, u_packed_ptr, h_ptr, y_ptr, stream),
                    "kml_generate_dev")

    def kmeans_dev(self, B, y_ptr, hhat_ptr, passes_ptr=0, stream=0):
        self._check(self._lib.kml_kmeans_dev(self._h, B, y_ptr, hhat_ptr, passes_ptr or None, stream), "kml_kmeans_dev")

    def receive_dev(self, B, y_ptr, var, uu_hat_packed_ptr, ret_ptr=0, true_h_ptr=0, stream=0):
        self._check(self._lib.kml_receive_dev(self._h, B, y_ptr, true_h_ptr or None, var, uu_hat_packed_ptr,
                                              ret_ptr or None, stream), "kml_receive_dev")

    def demap_dev(self, B, y_ptr, h_ptr, var, llr_ptr, stream=0):
        self._check(self._lib.kml_demap_dev(self._h, B, y_ptr, h_ptr, var, llr_ptr, stream), "kml_demap_dev")

    def decode_dev(self, B, llr_ptr, in_is_lr, iter_count, cc_hat_packed_ptr, ret_ptr, stream=0):
        self._check(self._lib.kml_decode_dev(self._h, B, llr_ptr, int(in_is_lr), iter_count, cc_hat_packed_ptr, ret_ptr,
                                             stream), "kml_decode_dev")

    def count_errors_dev(self, B, u_packed_ptr, uu_hat_packed_ptr, counters_ptr, stream=0):
        self._check(self._lib.kml_count_errors_dev(self._h, B, u_packed_ptr, uu_hat_packed_ptr, counters_ptr, stream),
                    "kml_count_errors_dev")


def snr_to_var(snr_db: float) -> float:
    """sigma^2 = 10^(-snr/10)  (src/simulator.cc:74-77)."""
    return math.pow(10.0, -0.1 * snr_db)


def unpack_bits(packed: np.ndarray, nbits: int) -> np.ndarray:
    """uint32 words (bit t = word t//32, bit t%32) → int8 [.., nbits]."""
    p = np.ascontiguousarray(packed, np.uint32)
    b = np.unpackbits(p.view(np.uint8), axis=-1, bitorder="little")
    return b[..., :nbits]


def pack_bits(bits: np.ndarray) -> np.ndarray:
    b = np.asarray(bits, np.uint8)
    n = b.shape[-1]
    pad = (-n) % 32
    if pad:
        b = np.concatenate([b, np.zeros(b.shape[:-1] + (pad,), np.uint8)], axis=-1)
    return np.packbits(b, axis=-1, bitorder="little").view(np.uint32)
