#!/usr/bin/env python
"""Build bev_b200/libbev_torch_ops.so -- the thin PyTorch C++ extension (csrc/torch_ops.cpp) that
registers torch.ops.bev_cuda.* on top of libbev_b200.so.  Plain g++ against the installed torch's
headers and libraries (no CUDA code in this file, so no nvcc); in-tree output, rpath $ORIGIN so the
extension finds libbev_b200.so next to it."""
import os
import subprocess
import sys

import torch
from torch.utils import cpp_extension

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libbev_torch_ops.so")
SRC = os.path.join(HERE, "torch_ops.cpp")


def build(force=False):
    deps = [SRC, os.path.join(PKG, "..", "include", "bev_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    inc = []
    for p in cpp_extension.include_paths("cuda"):
        inc += ["-isystem", p]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)] + inc + [
           SRC, "-o", OUT, "-L" + PKG, "-l:libbev_b200.so", "-L" + libdir, "-lc10", "-lc10_cuda",
           "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
