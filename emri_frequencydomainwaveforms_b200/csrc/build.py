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


HOST_SRC = os.path.join(HERE, "emrihost.c")
HOST_OUT = os.path.join(HERE, "libemrihost.so")


def build_host(force=False):
    """Native host-side producers (trajectory ODE, Schwarzschild frequencies): gcc + OpenMP, generic x86-64-v3."""
    if not force and os.path.exists(HOST_OUT) and os.path.getmtime(HOST_OUT) >= os.path.getmtime(HOST_SRC):
        return HOST_OUT
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-Wall", HOST_SRC, "-o", HOST_OUT, "-lm", "-lpthread"])
    return HOST_OUT


def build(force=False, verbose=False):
    build_host(force)
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared", "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
