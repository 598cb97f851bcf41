#!/bin/bash
# usage: gpu_r2_multi.sh N TAG   (under gpurun --gpus N)
set -u
N=${1:-2}; TAG=${2:-r2_n$N}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 600 $RUN scripts/p2p_check.py > gpurun_out/${TAG}_p2p_check.log 2>&1; echo "p2p_check exit $?"; grep p2p_check gpurun_out/${TAG}_p2p_check.log | tail -3
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
echo "bench cfg3 exit $?"; tail -2 gpurun_out/${TAG}_bench_cfg3.err; head -c 300 gpurun_out/${TAG}_bench_cfg3.json; echo
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 --config2-sharded --sustain-seconds 0 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
echo "bench cfg2 exit $?"; tail -2 gpurun_out/${TAG}_bench_cfg2.err; head -c 300 gpurun_out/${TAG}_bench_cfg2.json; echo
RB_SEG_BULK_PEER=1 timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 --config2-sharded --sustain-seconds 0 > gpurun_out/${TAG}_bench_cfg2_bulkpeer.json 2> gpurun_out/${TAG}_bench_cfg2_bulkpeer.err
echo "bench cfg2 bulk-peer exit $?"; tail -2 gpurun_out/${TAG}_bench_cfg2_bulkpeer.err; head -c 300 gpurun_out/${TAG}_bench_cfg2_bulkpeer.json; echo
