#!/bin/bash
cd "$(dirname "$0")/.."
W=cfg2_1080p_to_bev1024_u8c3_bilinear_x256
run() { echo "== $*"; env "$@" python tools/kbench.py --steps 30 $W; }
{
run A=0
run BEVK_PF=2
run BEVK_PF=4
run BEVK_SLACK=2
run BEVK_SLACK=1
run BEVK_SLACK=2 BEVK_PF=2
run BEVK_MAXFPS=8
run BEVK_MAXFPS=2
run BEVK_MAXFPS=2 BEVK_SLACK=2
cp bev_b200/libbev_b200_c4.so.keep bev_b200/libbev_b200.so
echo "#### 4 CTAs/SM build"
run A=0
run BEVK_SLACK=2
run BEVK_PF=2
run BEVK_DBG=3
cp bev_b200/libbev_b200_c3.so.keep bev_b200/libbev_b200.so
} > gpurun_out/r2_kb3.log 2>&1
cat gpurun_out/r2_kb3.log
