#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_warp_gpu.py -x -q -k "pixel_formats or hash or dtype_channel" > gpurun_out/r2_fmt_pytest.log 2>&1; echo rc=$? >> gpurun_out/r2_fmt_pytest.log
tail -12 gpurun_out/r2_fmt_pytest.log
python tools/kbench.py --steps 10 cfg2_u8c1_x256 cfg2_u8c4_x256 cfg2_f32c3_x128 > gpurun_out/r2_fmt_kb.log 2>&1
python tools/kbench.py --steps 10 --path generic cfg2_u8c1_x256 cfg2_u8c4_x256 cfg2_f32c3_x128 >> gpurun_out/r2_fmt_kb.log 2>&1
cat gpurun_out/r2_fmt_kb.log
