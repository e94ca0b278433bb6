"""Build tests/emul/libqr_emul.so: the product kernel sources compiled by g++ against the CUDA
emulation shim (cuda_emul.h).  TEST INFRASTRUCTURE ONLY -- see cuda_emul.h."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "qradient_b200", "csrc")
LIB = os.path.join(HERE, "libqr_emul.so")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("cuda_emul.h", "cuda_emul.cpp")]
    deps.append(os.path.join(ROOT, "include", "qradient_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not is_stale():
        return LIB
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DQR_HOST_EMUL", "-I", HERE, "-I", CSRC,
           "-x", "c++", os.path.join(CSRC, "qr_lib.cu"), os.path.join(HERE, "cuda_emul.cpp"), "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the emulation library")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
