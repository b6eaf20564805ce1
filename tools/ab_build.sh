#!/bin/bash
# Development aid: build a variant of the library (extra -D switches for warp_fast.cu) into
# tools/_ab/<name>.so for same-box A/B runs (tools/kbench.py --lib).   tools/ab_build.sh name -DX=1 ...
set -e
cd "$(dirname "$0")/../bev_b200/csrc"
name=$1; shift
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v --expt-relaxed-constexpr"
mkdir -p ../../tools/_ab
$NV "$@" -c warp_fast.cu -o ../../tools/_ab/$name.o 2> ../../tools/_ab/$name.ptxas.log || { cat ../../tools/_ab/$name.ptxas.log; exit 1; }
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_ab/$name.so ../../tools/_ab/$name.o capi.o warp_generic.o proj.o compo.o resize.o iou.o
grep -A1 "PxU8C3Lb1ELi6ELi4" ../../tools/_ab/$name.ptxas.log | grep -E "registers|spill" | head -3
