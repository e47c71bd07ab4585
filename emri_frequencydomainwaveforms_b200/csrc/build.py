"""Build libemrifd.so for sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emrifd.cu")
OUT = os.path.join(HERE, "libemrifd.so")
DEPS = [SRC, os.path.join(HERE, "k13_tables.h"), os.path.join(HERE, "..", "..", "include", "emrifd.h")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    mt = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > mt for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
