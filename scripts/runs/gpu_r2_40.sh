#!/bin/bash
set -u
N=2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 600 $RUN scripts/p2p_timeline.py --replicate-small 4096 > gpurun_out/r2_40_p2p_timeline_rep.txt 2>&1; echo "exit $?"; tail -3 gpurun_out/r2_40_p2p_timeline_rep.txt
