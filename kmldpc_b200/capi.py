"""ctypes binding of the C ABI declared in include/kmldpc_b200.h (the drop-in boundary).

No torch, no numpy conversions beyond pointer passing; no CPU fallback: if libkmldpc_b200.so is missing it is
built with nvcc (kmldpc_b200/build.py), and if that fails the import raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)


class KmlCode(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("n_graph", C.c_int32), ("n_tx", C.c_int32), ("k", C.c_int32),
                ("n_chk", C.c_int32), ("puncture", C.c_int32), ("info_offset", C.c_int32), ("n_edges", C.c_int32),
                ("is_5g", C.c_int32), ("encoder_active", C.c_int32), ("enc_words", C.c_int32), ("reserved", C.c_int32),
                ("row_ptr", c_i32p), ("col_idx", c_i32p), ("perm", c_i32p), ("enc_rows", c_u32p)]


class KmlModem(C.Structure):
    _fields_ = [("bits_per_symbol", C.c_int32), ("n_points", C.c_int32), ("points", c_f64p)]


class KmlOpts(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("known_h", C.c_int32), ("metric_type", C.c_int32),
                ("metric_iter", C.c_int32), ("kmeans_iter", C.c_int32), ("early_exit", C.c_int32),
                ("max_batch", C.c_int32), ("algorithm", C.c_int32)]


class KmlSweepCfg(C.Structure):
    _fields_ = [("min_snr", C.c_double), ("max_snr", C.c_double), ("step_snr", C.c_double),
                ("max_err_blk", C.c_uint64), ("max_num_blk", C.c_uint64),
                ("known_h", C.c_int32), ("is_5g", C.c_int32), ("metric_type", C.c_int32), ("metric_iter", C.c_int32),
                ("max_iter", C.c_int32), ("encoder_active", C.c_int32),
                ("histogram_enable", C.c_int32), ("reduce_on_host", C.c_int32),
                ("matrix_file", C.c_char * 512), ("modem_file", C.c_char * 512),
                ("seed", C.c_uint64),
                ("n_gpus", C.c_int32), ("max_batch", C.c_int32), ("early_exit", C.c_int32), ("algorithm", C.c_int32),
                ("debug_frames", C.c_int32), ("reserved2", C.c_int32)]


LOG_CB = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)

# name -> (restype, argtypes).  Every symbol include/kmldpc_b200.h declares is listed here; tests check the export table.
SYMBOLS = {
    "kml_code_load": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.POINTER(KmlCode))]),
    "kml_code_free": (None, [C.POINTER(KmlCode)]),
    "kml_modem_load": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(KmlModem))]),
    "kml_modem_free": (None, [C.POINTER(KmlModem)]),
    "kml_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(KmlCode), C.POINTER(KmlModem), C.POINTER(KmlOpts)]),
    "kml_destroy": (None, [C.c_void_p]),
    "kml_last_error": (C.c_char_p, [C.c_void_p]),
    "kml_set_early_exit": (C.c_int, [C.c_void_p, C.c_int]),
    "kml_set_algorithm": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "kml_set_minsum": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "kml_info": (C.c_int, [C.c_void_p, c_i32p]),
    "kml_decoder_info": (C.c_int, [C.c_void_p, c_i32p]),
    "kml_measure_smem_bandwidth": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "kml_launch_count": (C.c_uint64, [C.c_void_p]),
    "kml_encode": (C.c_int, [C.c_void_p, C.c_int, c_i32p, c_i32p]),
    "kml_generate": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64, c_i32p, c_i32p, c_f32p, c_f32p]),
    "kml_modulate": (C.c_int, [C.c_void_p, C.c_int, c_i32p, c_f32p, c_f32p, C.c_double, c_f32p]),
    "kml_kmeans": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, c_i32p]),
    "kml_kmeans_f64": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_f64p, c_i32p]),
    "kml_demap": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, C.c_double, c_f32p]),
    "kml_resolve": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, C.c_double, c_f32p, c_i32p]),
    "kml_decode": (C.c_int, [C.c_void_p, C.c_int, c_f32p, C.c_int, c_i32p, c_i32p, c_i32p]),
    "kml_decode_p0": (C.c_int, [C.c_void_p, C.c_int, c_f64p, C.c_int, c_i32p, c_i32p, c_i32p]),
    "kml_receive_f64": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_f64p, C.c_double, c_u32p, c_f64p, c_i32p, c_i32p, c_f32p]),
    "kml_soft_syndrome_state": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "kml_receive": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, C.c_double, c_u32p, c_f32p, c_i32p, c_i32p, c_f32p]),
    "kml_receive_submit": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, C.c_double, c_u32p, c_f32p, c_i32p, c_i32p, c_f32p]),
    "kml_receive_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "kml_count_errors": (C.c_int, [C.c_void_p, C.c_int, c_u32p, c_u32p, c_u64p]),
    "kml_simulate": (C.c_int, [C.c_void_p, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, c_u64p, c_u64p]),
    "kml_simulate_frames": (C.c_int, [C.c_void_p, C.c_double, C.c_uint64, C.c_uint64, C.c_int, c_u64p, c_f32p, c_f32p, c_f32p,
                                      c_i32p, c_i32p]),
    "kml_histogram": (C.c_int, [C.c_void_p, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, c_f32p, c_u64p]),
    "kml_histogram_rx": (C.c_int, [C.c_void_p, C.c_int, c_f32p, C.c_double, c_u32p, c_f32p, c_i32p, c_u32p, c_u64p]),
    "kml_generate_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "kml_kmeans_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kml_receive_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "kml_demap_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "kml_decode_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kml_count_errors_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kml_sweep_last_timing": (None, [c_f64p]),
    "kml_comm_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "kml_reduce_counters": (C.c_int, [C.c_void_p, c_u64p, C.c_int, c_u64p]),
    "kml_comm_destroy": (None, [C.c_void_p]),
    "kml_sweep_cfg_load": (C.c_int, [C.c_char_p, C.POINTER(KmlSweepCfg)]),
    "kml_sweep_points": (C.c_int, [C.POINTER(KmlSweepCfg)]),
    "kml_sweep_run": (C.c_int, [C.POINTER(KmlSweepCfg), C.c_char_p, c_f64p, c_f64p, c_u64p, LOG_CB, C.c_void_p]),
}

_lib = None


def load(rebuild: bool = False) -> C.CDLL:
    """dlopen the in-tree library (building it first if sources are newer).  Raises if it cannot be had."""
    global _lib
    if _lib is None or rebuild:
        path = _build.build(force=rebuild) if (rebuild or _build.is_stale()) else _build.LIB_PATH
        lib = C.CDLL(path, mode=getattr(os, "RTLD_NOW", 2))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError = missing export: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
