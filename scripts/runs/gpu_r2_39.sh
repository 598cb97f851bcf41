#!/bin/bash
# usage: gpu_r2_39.sh N   (under gpurun --gpus N)
set -u
N=${1:-2}; TAG=r2_39_n$N
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 600 $RUN scripts/p2p_check.py > gpurun_out/${TAG}_p2p_check.log 2>&1; echo "p2p_check exit $?"; grep -E "p2p_check|Error|error" gpurun_out/${TAG}_p2p_check.log | tail -6 | cut -c1-260
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 --replicate-small 4096 --sustain-seconds 0 > gpurun_out/${TAG}_bench_cfg3_rep.json 2> gpurun_out/${TAG}_bench_cfg3_rep.err
echo "bench cfg3 rep exit $?"; tail -3 gpurun_out/${TAG}_bench_cfg3_rep.err | cut -c1-300; head -c 260 gpurun_out/${TAG}_bench_cfg3_rep.json; echo
