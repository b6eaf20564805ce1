#!/bin/bash
# A/B on one box: round-1 tree (tools/_r1, built from commit 5571169) vs the current one, kernel-only timings
cd "$(dirname "$0")/.."
W="cfg2_1080p_to_bev1024_u8c3_bilinear_x256 cfg2_nearest cfg5_4k_to_bev2048_u8c3_x64 cfg5_inv_bev2048_to_4k_u8c3_x64 cfg5_4k_to_bev2048_f16c3_x64 cfg5_inv_bev2048_to_4k_f16c3_x64"
{
echo "== round-1 library"; (cd tools/_r1 && python tools/kbench.py --steps 30 $W)
echo "== current library"; python tools/kbench.py --steps 30 --check $W
} > gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log
