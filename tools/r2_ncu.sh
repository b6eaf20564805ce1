#!/bin/bash
# ncu --set full of the staged warp kernel on cfg 2 (after the same command ran clean without ncu)
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-projection --no-other-configs"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:warp_fast -s 4 -c 1 -o gpurun_out/r2_warp_${1:-a} -f $CMD > gpurun_out/r2_ncu_${1:-a}.log 2>&1
echo rc=$?; tail -3 gpurun_out/r2_ncu_plain.log | cut -c1-600
