#!/bin/bash
set -u
N=${1:-2}; TAG=r2_52_n$N
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 --dist zipf --sustain-seconds 0 > gpurun_out/${TAG}_bench_cfg3_zipf.json 2> gpurun_out/${TAG}_bench_cfg3_zipf.err
echo "bench cfg3 zipf exit $?"; tail -2 gpurun_out/${TAG}_bench_cfg3_zipf.err | cut -c1-200; grep "^{" gpurun_out/${TAG}_bench_cfg3_zipf.json | head -c 260; echo
