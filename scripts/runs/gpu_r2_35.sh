#!/bin/bash
# usage: gpu_r2_35.sh N   (under gpurun --gpus N)
set -u
N=${1:-2}; TAG=r2_35_n$N
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
echo "bench cfg3 exit $?"; tail -2 gpurun_out/${TAG}_bench_cfg3.err; head -c 260 gpurun_out/${TAG}_bench_cfg3.json; echo
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 --config2-sharded --sustain-seconds 0 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
echo "bench cfg2 exit $?"; tail -2 gpurun_out/${TAG}_bench_cfg2.err; head -c 260 gpurun_out/${TAG}_bench_cfg2.json; echo
