#!/bin/bash
cd "$(dirname "$0")/.."
W=cfg2_1080p_to_bev1024_u8c3_bilinear_x256
run() { echo "== $*"; env "$@" python tools/kbench.py --steps 30 --check $W; }
{
run A=0
run BEVK_YGROUP=1
run BEVK_YGROUP=2
run BEVK_YGROUP=4
run BEVK_YGROUP=8
run BEVK_YGROUP=1 BEVK_DBG=3
} > gpurun_out/r2_kb5.log 2>&1
grep "==" -A1 gpurun_out/r2_kb5.log | grep -v "^--" | paste - - | cut -c1-180
