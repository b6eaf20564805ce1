#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
$TR --master-port 29521 tools/gather_probe.py
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 $TR --master-port 29522 tools/gather_probe.py
NCCL_MIN_P2P_NCHANNELS=64 NCCL_MAX_P2P_NCHANNELS=64 NCCL_MAX_NCHANNELS=64 $TR --master-port 29523 tools/gather_probe.py
NCCL_P2P_NET_CHUNKSIZE=4194304 NCCL_BUFFSIZE=16777216 $TR --master-port 29524 tools/gather_probe.py
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_BUFFSIZE=16777216 $TR --master-port 29525 tools/gather_probe.py
} 2>&1 | grep "^N=" > gpurun_out/r2_gather_n$N.log
cat gpurun_out/r2_gather_n$N.log
