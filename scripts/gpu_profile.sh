#!/bin/bash
# One gpurun call: plain runs first (must exit 0), then the ncu launch list of the eager bench and one
# `--set full` capture of each hot kernel from the micro-benchmark.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-rX}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph"
KB="python scripts/kbench.py --ops fwd,bwd,update --iters 2"
timeout 600 $CMD > gpurun_out/plain_bench.log 2>&1 && timeout 300 $KB > gpurun_out/plain_kbench.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_bench.log gpurun_out/plain_kbench.log; exit 1; }
tail -1 gpurun_out/plain_kbench.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit: $?"
for K in dot_interaction_fwd_kernel dot_interaction_bwd_kernel seg_reduce_tiles_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_$K $KB > gpurun_out/ncu_$K.log 2>&1
  echo "ncu $K exit: $?"
done
ls -la gpurun_out/*.ncu-rep
