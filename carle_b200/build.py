"""Build recipe for libcarle_b200.so (in-tree, sm_100a only).

    python -m carle_b200.build [--verbose]

The shared library has no torch dependency: plain nvcc, static cudart, C ABI
(include/carle_b200.h).  It is git-ignored but travels with the tree to the GPU box.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "carle_abi.cu")
DEPS = [SRC, os.path.join(PKG, "csrc", "kernels.cuh"), os.path.join(PKG, "csrc", "ca_core.cuh"),
        os.path.join(ROOT, "include", "carle_b200.h")]
OUT = os.path.join(PKG, "lib", "libcarle_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    extra = os.environ.get("CARLE_NVCC_EXTRA", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", OUT, SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libcarle_b200.so:\n" + " ".join(cmd))
    return OUT


if __name__ == "__main__":
    build(force=True, verbose="--verbose" in sys.argv)
    print(OUT)
