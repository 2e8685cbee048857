"""Builds liblr_b200.so (sm_100a) in-tree with nvcc. Cross-compiles without a GPU.

    python -m leftrefill_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblr_b200.so")
SOURCES = ["ops.cu", "engine.cu"]
HEADERS = ["ptx.cuh", "gemm_tc.cuh", "attention_tc.cuh", "elementwise.cuh", "ops.h",
           os.path.join("..", "..", "include", "lr_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    if not force and not is_stale():
        return LIB
    tmp = f"{LIB}.{os.getpid()}.tmp"  # written elsewhere and renamed: other processes never see a partial library
    cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]
    if verbose:
        print("[leftrefill_b200.build]", " ".join(cmd), flush=True)
    try:
        subprocess.run(cmd, check=True)
        os.replace(tmp, LIB)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
