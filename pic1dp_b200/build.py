"""Builds libpic1dp_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpic1dp_b200.so")
SOURCES = ["pic1dp_gpu.cu"]
HEADERS = ["particle_kernels.cuh", "field_kernels.cuh", "diag_kernels.cuh", "optimize_kernels.cuh", "optimize_host.hpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",   # host-side arithmetic (marker optimisation) must not be contracted to FMA
]


def _source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to a GPU box)."""
    import hashlib
    h = hashlib.sha256()
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(HERE, "..", "include", "pic1dp_gpu.h"))
    for d in deps:
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".hash"):
        return True
    return open(LIB + ".hash").read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or needs_build()):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    with open(LIB + ".hash", "w") as f:
        f.write(_source_hash())
    return LIB


HOST_SRC = os.path.join(HERE, "..", "host", "pic1dp_host.cpp")
HOST_EXE = os.path.join(HERE, "..", "host", "pic1dp_host")


def build_host(force: bool = False) -> str:
    """C++ host driver above the C ABI (host/pic1dp_host.cpp): the reference's `program pic1dp` sequence."""
    build()
    import hashlib
    hh = hashlib.sha256(open(HOST_SRC, "rb").read() + open(os.path.join(HERE, "..", "include", "pic1dp_gpu.h"), "rb").read()).hexdigest()
    if not force and os.path.exists(HOST_EXE) and os.path.exists(HOST_EXE + ".hash") and \
            open(HOST_EXE + ".hash").read().strip() == hh:
        return HOST_EXE
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(HERE, "..", "include"), HOST_SRC,
           "-L", HERE, "-lpic1dp_b200", "-Wl,-rpath,$ORIGIN/../pic1dp_b200", "-o", HOST_EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    with open(HOST_EXE + ".hash", "w") as f:
        f.write(hh)
    return HOST_EXE


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
