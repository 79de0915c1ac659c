"""Builds dexdeform_b200/libmaniskill_mpm.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libmaniskill_mpm.so")
SOURCES = ["abi1_kernels.cu", "engine.cu", "fk.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-deprecated-declarations", "-shared"]


def _stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "dexdeform_mpm.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=True, extra=()):
    if not force and not _stale():
        return OUT
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = ["nvcc"] + flags + list(extra) + srcs + ["-o", OUT]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force=True, extra=sys.argv[1:])
