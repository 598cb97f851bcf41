#!/bin/bash
# usage: gpu_r2_41.sh N   (under gpurun --gpus N): the driver's N > 1 command
set -u
N=${1:-2}; TAG=r2_41_n$N
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench.err | cut -c1-200; head -c 260 gpurun_out/${TAG}_bench.json; echo
