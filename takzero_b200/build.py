"""Builds takzero_b200/libtakzero_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

The library is the product: there is no CPU fallback, so importing the package on a GPU box
without this file fails loudly."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtakzero_b200.so")
SOURCES = ["api.cu", "kernels.cu", "nn.cu", "rnd.cu", "comm.cu", "model_file.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "static",
]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hostdir = os.path.join(HERE, "..", "host")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(hostdir, f) for f in os.listdir(hostdir)] + [
        os.path.join(HERE, "..", "include", "takzero_b200.h"), os.path.join(HERE, "..", "include", "takzero_b200.hpp"),
        os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose:
            print(out)
    cmd = [nvcc(), "-shared", "-o", LIB, *objs, "--cudart", "static",
           "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    subprocess.check_call(cmd)
    build_hosts()
    return LIB


HOSTS = ["selfplay", "reanalyze", "evaluation", "tei", "analysis", "format_check"]


def build_hosts() -> None:
    """C++ host programs over the C ABI (host/*.cpp) -> takzero_b200/bin/."""
    bindir = os.path.join(HERE, "bin")
    os.makedirs(bindir, exist_ok=True)
    for name in HOSTS:
        src = os.path.join(HERE, "..", "host", f"{name}.cpp")
        if not os.path.exists(src):
            continue
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", src, "-o", os.path.join(bindir, name),
                               f"-L{HERE}", "-ltakzero_b200", "-Wl,-rpath,$ORIGIN/.."])


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
