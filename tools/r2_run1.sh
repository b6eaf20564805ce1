#!/bin/bash
# round-2 development run: warp parity tests, then kernel-only timings with the new paths toggled
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_warp_gpu.py -x -q > gpurun_out/r2_pytest_warp1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_warp1.log
tail -5 gpurun_out/r2_pytest_warp1.log
{
echo "== default"; timeout 300 python tools/kbench.py --check cfg2_1080p_to_bev1024_u8c3_bilinear_x256 cfg2_nearest cfg5_4k_to_bev2048_u8c3_x64 cfg5_inv_bev2048_to_4k_u8c3_x64 cfg5_4k_to_bev2048_f16c3_x64
echo "== no setup cache"; BEVK_NO_SETUP_CACHE=1 timeout 300 python tools/kbench.py cfg2_1080p_to_bev1024_u8c3_bilinear_x256 cfg2_nearest
echo "== no pairs"; BEVK_NO_PAIRS=1 timeout 300 python tools/kbench.py cfg2_1080p_to_bev1024_u8c3_bilinear_x256
echo "== segs 2"; BEVK_FAST_SEGS=2 timeout 300 python tools/kbench.py cfg2_1080p_to_bev1024_u8c3_bilinear_x256
echo "== segs 1"; BEVK_FAST_SEGS=1 timeout 300 python tools/kbench.py cfg2_1080p_to_bev1024_u8c3_bilinear_x256
} > gpurun_out/r2_kb1.log 2>&1
cat gpurun_out/r2_kb1.log
timeout 120 compute-sanitizer --tool racecheck python -c "print('hello')" > gpurun_out/r2_sanitizer_probe.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/r2_sanitizer_probe.log
tail -3 gpurun_out/r2_sanitizer_probe.log
