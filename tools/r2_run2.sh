#!/bin/bash
cd "$(dirname "$0")/.."
W=cfg2_1080p_to_bev1024_u8c3_bilinear_x256
{
echo "== default"; python tools/kbench.py $W
echo "== dbg1 no stores"; BEVK_DBG=1 python tools/kbench.py $W
echo "== dbg2 no tma"; BEVK_DBG=2 python tools/kbench.py $W
echo "== dbg3 neither"; BEVK_DBG=3 python tools/kbench.py $W
} > gpurun_out/r2_kb2.log 2>&1
cat gpurun_out/r2_kb2.log
