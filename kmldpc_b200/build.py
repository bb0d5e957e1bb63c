"""Builds kmldpc_b200/lib/libkmldpc_b200.so IN-TREE with nvcc for sm_100a only (no other arch, no JIT cache)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libkmldpc_b200.so")
SOURCES = ["bp_decode.cu", "bp_minsum.cu", "bp_layered.cu", "link_kernels.cu", "kml_api.cu", "host_code.cpp", "layout_opt.cpp", "sweep.cpp"]
HEADERS = ["kml_internal.h", "kml_kernels.cuh", "bp_minsum_nodes.cuh", os.path.join("..", "..", "include", "kmldpc_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function"]
LINK_FLAGS = ["--shared", "-lpthread", "-ldl"]
OBJ_DIR = os.path.join(HERE, "lib", "obj")


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, tuning: bool = False) -> str:
    """One nvcc -c per source (in parallel, only the stale ones), then one link.  tuning=True adds -DKML_TUNING (the A/B
    and timing-ablation kernel variants, kml_internal.h) — never for the shipped library."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DKML_TUNING"] if tuning else [])
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src + (".tuning.o" if tuning else ".o"))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(os.path.join(CSRC, src))):
            subprocess.check_call([nvcc()] + flags + ["-c", os.path.join(CSRC, src), "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc()] + NVCC_FLAGS + LINK_FLAGS + objs + ["-o", LIB_PATH])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tuning="--tuning" in sys.argv))
