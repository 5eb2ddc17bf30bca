"""Build libnadavca_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

Usage: ``python -m nadavca_b200.build [--force]``.  nvcc cross-compiles without a GPU.  The .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnadavca_b200.so")
SOURCES = ["api.cu", "band.cu", "rows4.cu", "rows5.cu", "snp3.cu", "path2.cu", "finalize.cu", "select.cu", "anchors.cu", "microbench.cu"]
HEADERS = ["common.cuh", "dp3.cuh", "kernels.h", os.path.join("..", "..", "include", "nadavca_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # the log-space kernels round like the reference's scalar C++ (no FMA contraction); kernels that want FMA call
    # fma() explicitly
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "0",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("NVB_EXTRA_FLAGS", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libnadavca_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
