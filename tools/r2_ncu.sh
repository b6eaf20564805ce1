#!/bin/bash
# round 2: ncu evidence for the staged warp kernel and the launch list of the bench command
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs"
$CMD > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo launches rc=$?
CMD2="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-projection --no-other-configs"
$CMD2 > gpurun_out/r02_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:warp_fast -s 4 -c 2 -o gpurun_out/r02_warp -f $CMD2 > gpurun_out/r02_ncu_full.log 2>&1
echo full rc=$?
