#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -15 gpurun_out/r2_pytest_gpu.log
