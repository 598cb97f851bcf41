#!/bin/bash
set -u
N=2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
for w in 0 8 5; do
RB_IX16_WARPS=$w timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 --config2-sharded --sustain-seconds 0 --no-e2e > gpurun_out/r2_14_n2_cfg2_w$w.json 2> gpurun_out/r2_14_n2_cfg2_w$w.err
echo "warps=$w exit $?"; grep '^{' gpurun_out/r2_14_n2_cfg2_w$w.json | head -c 220; echo
done
