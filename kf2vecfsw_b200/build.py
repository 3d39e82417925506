"""In-tree build of libkfcount.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "csrc", "kf_api.cu"), os.path.join(HERE, "csrc", "kf_host.cpp")]
HEADERS = [os.path.join(HERE, "csrc", "kf_kernels.cuh"), os.path.join(HERE, "csrc", "kf_sparse.cuh"),
           os.path.join(HERE, "csrc", "kf_sparse_host.inc"), os.path.join(ROOT, "include", "kfcount.h")]
OUT = os.path.join(HERE, "libkfcount.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("KF_NVCC_EXTRA", "").split()   # developer experiments (-DKF_... switches)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + SOURCES + ["-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
