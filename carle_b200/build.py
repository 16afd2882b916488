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
CSRC = os.path.join(PKG, "csrc")
# translation units (compiled in parallel, then linked into one shared library)
SOURCES = [os.path.join(CSRC, name) for name in
           ("carle_abi.cu", "stream_abi.cu", "fused_abi.cu", "strip_abi.cu", "random_abi.cu", "wrappers_abi.cu", "host_pack.cu",
            "jit.cu")]
# kernel headers embedded into the library for run-time (NVRTC) rule specialisation, jit.cu
EMBEDDED = [("kSrcCaCore", "ca_core.cuh"), ("kSrcKernels", "kernels.cuh"), ("kSrcStrip", "strip.cuh"),
            ("kSrcTiled", "tiled.cuh")]
OUT = os.path.join(PKG, "lib", "libcarle_b200.so")
OBJ_DIR = os.path.join(PKG, "lib", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def deps():
    found = [os.path.join(ROOT, "include", "carle_b200.h"), os.path.abspath(__file__)]
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            found.append(os.path.join(CSRC, name))
    return found


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def source_hash():
    """sha256 over the library's sources (names + contents): what the built .so is checked
    against -- file times do not survive the copy to the GPU box."""
    import hashlib
    h = hashlib.sha256()
    for d in deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def up_to_date():
    stamp = OUT + ".srchash"
    if not (os.path.exists(OUT) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == source_hash()


def build(force=False, verbose=False, out=None):
    """Compile every translation unit for sm_100a and link libcarle_b200.so.
    `out` (with CARLE_NVCC_EXTRA) builds an alternative library for A/B runs
    (select it at run time with CARLE_B200_LIB)."""
    out = out or OUT
    if not force and out == OUT and up_to_date():
        return out
    stamp = source_hash()          # (of the sources as they are NOW: an edit during the build must not pass as built)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tag = os.path.splitext(os.path.basename(out))[0]
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = os.environ.get("CARLE_NVCC_EXTRA", "").split()
    nvcc = nvcc_path()
    inc_dir = os.path.join(OBJ_DIR, tag + ".gen")
    os.makedirs(inc_dir, exist_ok=True)
    with open(os.path.join(inc_dir, "embedded_sources.inc"), "w") as f:
        for symbol, name in EMBEDDED:
            text = open(os.path.join(CSRC, name)).read()
            assert ')CARLE_SRC"' not in text
            # (adjacent raw literals: some compilers cap a single literal at 64 KiB)
            f.write("static const char %s[] =\n" % symbol)
            for i in range(0, len(text), 16000):
                f.write('R"CARLE_SRC(' + text[i:i + 16000] + ')CARLE_SRC"\n')
            f.write(";\n")
    extra = extra + ["-I", inc_dir]
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, tag + "." + os.path.splitext(os.path.basename(src))[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", "-o", obj, src]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                 stderr=subprocess.STDOUT, text=True)))
    objs = []
    for cmd, obj, proc in procs:
        log = proc.communicate()[0]
        if verbose or proc.returncode != 0:
            sys.stderr.write(log)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed building libcarle_b200.so:\n" + " ".join(cmd))
        objs.append(obj)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
            "-o", out] + objs + ["-ldl"]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("link failed building libcarle_b200.so:\n" + " ".join(link))
    if out == OUT and not os.environ.get("CARLE_NVCC_EXTRA"):
        with open(OUT + ".srchash", "w") as f:
            f.write(stamp + "\n")
    # the intermediates (35 MB per build) are of no use once linked, and everything in the tree
    # travels to the GPU box
    import shutil
    for obj in objs:
        os.remove(obj)
    shutil.rmtree(inc_dir, ignore_errors=True)
    return out


if __name__ == "__main__":
    target = None
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            target = os.path.abspath(a[6:])
    print(build(force=True, verbose="--verbose" in sys.argv, out=target))
