#!/bin/bash
# round-2: the bench as the driver runs it (N = $1), ours then the reference arm
cd "$(dirname "$0")/.."
N=${1:-1}
if [ "$N" = "1" ]; then
  python bench.py --impl reference --gpus 1 --steps 20 --warmup 2 > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/r2_bench_ref_n1.err
  python bench.py --gpus 1 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
fi
echo rc=$?
tail -c 3000 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
