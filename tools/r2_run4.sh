#!/bin/bash
cd "$(dirname "$0")/.."
W=cfg2_1080p_to_bev1024_u8c3_bilinear_x256
run() { echo "== $*"; env "$@" python tools/kbench.py --steps 30 $W; }
{
run A=0
run BEVK_DBG=4
run BEVK_DBG=4 BEVK_SLACK=1
run BEVK_DBG=4 BEVK_MAXFPS=8
run BEVK_DBG=4 BEVK_FAST_SEGS=2
run BEVK_DBG=4 BEVK_FAST_SEGS=1
} > gpurun_out/r2_kb4.log 2>&1
grep "==" -A1 gpurun_out/r2_kb4.log | grep -v "^--" | paste - - | cut -c1-160
