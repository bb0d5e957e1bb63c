"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the CPU oracle (oracle/kml_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker.  kmldpc_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkml_oracle.so")
CONFIG_DIR = os.path.join(os.path.dirname(_HERE), "config")
REF_HARNESS = os.path.join(_HERE, "_ref", "ref_harness")


def build(force: bool = False) -> str:
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(_HERE, "kml_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "kml_oracle.h")))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "libkml_oracle.so")])
    return _LIB_PATH


class _Opts(C.Structure):
    _fields_ = [("known_h", C.c_int), ("is_5g", C.c_int), ("metric_type", C.c_int), ("metric_iter", C.c_int),
                ("max_iter", C.c_int), ("kmeans_iter", C.c_int), ("histogram", C.c_int)]


class _FrameOut(C.Structure):
    _fields_ = [("h", C.c_double * 2), ("hhat", C.c_double * 2), ("metric", C.c_double * 4),
                ("kstar", C.c_int), ("ret", C.c_int), ("nerr", C.c_int)]


class _Lcg(C.Structure):
    _fields_ = [("state", C.c_long)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.kmo_code_load.restype = C.c_void_p
        L.kmo_code_load.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.kmo_modem_load.restype = C.c_void_p
        L.kmo_modem_load.argtypes = [C.c_char_p]
        L.kmo_lcg_uniform.restype = C.c_double
        L.kmo_run.restype = C.c_int64
        for name in ("kmo_code_free", "kmo_modem_free"):
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


@dataclass
class FrameResult:
    h: complex
    hhat: complex
    metric: np.ndarray
    kstar: int
    ret: int
    nerr: int
    u: np.ndarray | None = None
    c: np.ndarray | None = None
    y: np.ndarray | None = None
    clusters: np.ndarray | None = None
    p0: np.ndarray | None = None
    cc_hat: np.ndarray | None = None
    uu_hat: np.ndarray | None = None


class Lcg:
    """lib/lab/src/randnum.cc CLCRandNum; seed(-1) ⇒ state 17."""

    def __init__(self, state: int = 17):
        self._g = _Lcg(state)

    def uniform(self) -> float:
        return lib().kmo_lcg_uniform(C.byref(self._g))

    def normal(self, n: int) -> np.ndarray:
        out = np.empty(n, np.float64)
        lib().kmo_lcg_normal(C.byref(self._g), _p(out, C.c_double), C.c_int(n))
        return out

    @property
    def state(self) -> int:
        return self._g.state


class Code:
    def __init__(self, h_file: str, is_5g: bool = False, active: bool = True):
        path = h_file if os.path.isabs(h_file) else os.path.join(CONFIG_DIR, h_file)
        self._h = lib().kmo_code_load(path.encode(), int(is_5g), int(active))
        if not self._h:
            raise FileNotFoundError(path)
        info = (C.c_int32 * 8)()
        lib().kmo_code_info(C.c_void_p(self._h), info)
        (self.M, self.N, self.N_tx, self.K, self.chk, self.two_z, self.E, act) = list(info)
        self.active = bool(act)
        self.is_5g = bool(is_5g)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kmo_code_free(C.c_void_p(self._h))
            self._h = None

    def export(self, with_enc: bool = True):
        row_ptr = np.empty(self.M + 1, np.int32)
        col_idx = np.empty(self.E, np.int32)
        col_ptr = np.empty(self.N + 1, np.int32)
        row_idx = np.empty(self.E, np.int32)
        perm = np.empty(self.N, np.int32)
        enc = np.zeros((self.M, self.N), np.uint8) if (with_enc and self.active) else None
        lib().kmo_code_export(C.c_void_p(self._h), _p(row_ptr, C.c_int32), _p(col_idx, C.c_int32),
                              _p(col_ptr, C.c_int32), _p(row_idx, C.c_int32), _p(perm, C.c_int32),
                              _p(enc, C.c_uint8))
        return dict(row_ptr=row_ptr, col_idx=col_idx, col_ptr=col_ptr, row_idx=row_idx, perm=perm, enc_h=enc)

    def encode(self, u: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(u, np.int32).copy()
        c = np.empty(self.N_tx, np.int32)
        lib().kmo_encode(C.c_void_p(self._h), _p(u, C.c_int), _p(c, C.c_int))
        return c

    def parity_check(self, rr: np.ndarray) -> int:
        rr = np.ascontiguousarray(rr, np.int32)
        return lib().kmo_parity_check(C.c_void_p(self._h), _p(rr, C.c_int))

    def decode(self, p0: np.ndarray, iter_count: int, max_iter: int | None = None):
        p0 = np.ascontiguousarray(p0, np.float64)
        uu = np.empty(self.K, np.int32)
        cc = np.empty(self.N, np.int32)
        soft = np.ones(self.M, np.float64)
        ret = lib().kmo_decode(C.c_void_p(self._h), _p(p0, C.c_double), C.c_int(iter_count),
                               C.c_int(iter_count if max_iter is None else max_iter), _p(uu, C.c_int),
                               _p(cc, C.c_int), _p(soft, C.c_double))
        return ret, uu, cc, soft


class Modem:
    def __init__(self, modem_file: str):
        path = modem_file if os.path.isabs(modem_file) else os.path.join(CONFIG_DIR, modem_file)
        self._h = lib().kmo_modem_load(path.encode())
        if not self._h:
            raise FileNotFoundError(path)
        info = (C.c_int32 * 2)()
        lib().kmo_modem_info(C.c_void_p(self._h), info)
        self.bits, self.Q = info[0], info[1]
        pts = np.empty(2 * self.Q, np.float64)
        lib().kmo_modem_points(C.c_void_p(self._h), _p(pts, C.c_double))
        self.points = pts.view(np.complex128)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kmo_modem_free(C.c_void_p(self._h))
            self._h = None

    def map(self, c: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(c, np.int32)
        n = c.size // self.bits
        xx = np.empty(2 * n, np.float64)
        lib().kmo_map(C.c_void_p(self._h), _p(c, C.c_int), C.c_int(n), _p(xx, C.c_double))
        return xx.view(np.complex128)

    def demap(self, y: np.ndarray, h: complex, var: float) -> np.ndarray:
        y = np.ascontiguousarray(y, np.complex128)
        p0 = np.empty(y.size * self.bits, np.float64)
        lib().kmo_demap(C.c_void_p(self._h), _p(y.view(np.float64), C.c_double), C.c_int(y.size),
                        C.c_double(h.real), C.c_double(h.imag), C.c_double(var), _p(p0, C.c_double))
        return p0


def kmeans(y: np.ndarray, points: np.ndarray, iters: int = 20):
    y = np.ascontiguousarray(y, np.complex128)
    pts = np.ascontiguousarray(points, np.complex128)
    cl = np.empty(pts.size, np.complex128)
    passes = lib().kmo_kmeans(_p(y.view(np.float64), C.c_double), C.c_int(y.size),
                              _p(pts.view(np.float64), C.c_double), C.c_int(pts.size), C.c_int(iters),
                              _p(cl.view(np.float64), C.c_double))
    return cl, passes


class Link:
    """One (code, modem, options) triple — the oracle's view of Simulator + KmCodec."""

    def __init__(self, matrix: str, modem: str, *, is_5g=False, active=True, known_h=False, metric_type=False,
                 metric_iter=5, max_iter=50, kmeans_iter=20, histogram=False):
        self.code = Code(matrix, is_5g, active)
        self.modem = Modem(modem)
        self.opts = _Opts(int(known_h), int(is_5g), int(metric_type), metric_iter, max_iter, kmeans_iter, int(histogram))
        self.n_sym = self.code.N_tx // self.modem.bits
        # the codec's syndrom_soft_ member: survives from frame to frame (binaryldpccodec.cc:274 is its only writer)
        self.soft_state = np.ones(self.code.M, np.float64)

    def reset_soft_state(self):
        self.soft_state[:] = 1.0

    def frame(self, lcg: Lcg, snr_db: float, full: bool = True) -> FrameResult:
        c, m = self.code, self.modem
        fo = _FrameOut()
        if full:
            u = np.empty(c.K, np.int32); cw = np.empty(c.N_tx, np.int32); y = np.empty(2 * self.n_sym, np.float64)
            cl = np.empty(2 * m.Q, np.float64); p0 = np.empty(c.N_tx, np.float64)
            cch = np.empty(c.N, np.int32); uh = np.empty(c.K, np.int32)
        else:
            u = cw = y = cl = p0 = cch = uh = None
        lib().kmo_frame(C.c_void_p(c._h), C.c_void_p(m._h), C.byref(self.opts), C.byref(lcg._g), C.c_double(snr_db),
                        C.byref(fo), _p(u, C.c_int), _p(cw, C.c_int), _p(y, C.c_double), _p(cl, C.c_double),
                        _p(p0, C.c_double), _p(cch, C.c_int), _p(uh, C.c_int), _p(self.soft_state, C.c_double))
        return FrameResult(complex(*fo.h), complex(*fo.hhat), np.array(fo.metric), fo.kstar, fo.ret, fo.nerr,
                           u, cw, None if y is None else y.view(np.complex128),
                           None if cl is None else cl.view(np.complex128), p0, cch, uh)

    def receive(self, y: np.ndarray, var: float, true_h: complex = 0j) -> FrameResult:
        c, m = self.code, self.modem
        y = np.ascontiguousarray(y, np.complex128)
        fo = _FrameOut()
        th = np.array([true_h.real, true_h.imag], np.float64)
        cl = np.empty(2 * m.Q, np.float64); p0 = np.empty(c.N_tx, np.float64)
        cch = np.empty(c.N, np.int32); uh = np.empty(c.K, np.int32)
        lib().kmo_receive(C.c_void_p(c._h), C.c_void_p(m._h), C.byref(self.opts), _p(y.view(np.float64), C.c_double),
                          _p(th, C.c_double), C.c_double(var), C.byref(fo), _p(cl, C.c_double), _p(p0, C.c_double),
                          _p(cch, C.c_int), _p(uh, C.c_int), _p(self.soft_state, C.c_double))
        return FrameResult(true_h, complex(*fo.hhat), np.array(fo.metric), fo.kstar, fo.ret, -1,
                           None, None, y, cl.view(np.complex128), p0, cch, uh)

    def bulk(self, snr_db: float, frames: int, threads: int | None = None, frame0: int = 0, chain_block: int = 1) -> dict:
        """`frames` reference frames (frame f has its own LCG), computed on `threads` host threads.  Blocks of
        `chain_block` consecutive frames share one syndrom_soft_ state in frame order (soft metric only)."""
        c, m = self.code, self.modem
        threads = threads or (os.cpu_count() or 1)
        out = dict(y=np.empty((frames, self.n_sym, 2), np.float64), h=np.empty((frames, 2), np.float64),
                   hhat=np.empty((frames, 2), np.float64), kstar=np.empty(frames, np.int32), ret=np.empty(frames, np.int32),
                   nerr=np.empty(frames, np.int32), converged=np.empty(frames, np.uint8),
                   u=np.empty((frames, c.K), np.uint8), uu_hat=np.empty((frames, c.K), np.uint8),
                   metric=np.zeros((frames, 4), np.float64))
        lib().kmo_bulk(C.c_void_p(c._h), C.c_void_p(m._h), C.byref(self.opts), C.c_double(snr_db), C.c_long(frame0),
                       C.c_long(frames), C.c_int(threads), C.c_long(chain_block), _p(out["y"], C.c_double), _p(out["h"], C.c_double),
                       _p(out["hhat"], C.c_double), _p(out["kstar"], C.c_int32), _p(out["ret"], C.c_int32),
                       _p(out["nerr"], C.c_int32), _p(out["converged"], C.c_uint8), _p(out["u"], C.c_uint8),
                       _p(out["uu_hat"], C.c_uint8), _p(out["metric"], C.c_double))
        out["y"] = out["y"].view(np.complex128).reshape(frames, self.n_sym)
        out["h"] = out["h"].view(np.complex128).reshape(frames)
        out["hhat"] = out["hhat"].view(np.complex128).reshape(frames)
        return out

    def run(self, snr_db: float, frames: int, threads: int = 1, seed0: int = 17):
        cnt = (C.c_uint64 * 4)()
        iters = lib().kmo_run(C.c_void_p(self.code._h), C.c_void_p(self.modem._h), C.byref(self.opts),
                              C.c_double(snr_db), C.c_long(seed0), C.c_long(frames), C.c_int(threads), cnt)
        return list(cnt), int(iters)
