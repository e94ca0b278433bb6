"""Build libqradient_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch dependency).

    python -m qradient_b200.build            # build if stale
    python -m qradient_b200.build --force

The library is a single translation unit (csrc/qr_lib.cu) with the CUDA runtime linked statically,
so it loads through ctypes without LD_LIBRARY_PATH and without torch.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libqradient_b200.so")
SOURCES = ["qr_lib.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "qradient_b200.h")]   # every header of the single translation unit

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--cudart", "static",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libqradient_b200.so")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra=()):
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + [os.path.join(CSRC, f) for f in SOURCES] + ["-o", LIB]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libqradient_b200.so")
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    extra = []
    if "--ptxas-v" in sys.argv:
        extra += ["-Xptxas", "-v"]
    print(build(force="--force" in sys.argv, verbose=True, extra=extra))
