#!/bin/bash
# Round-2 evidence in one gpurun call: plain runs first (must exit 0), then the ncu launch list of the eager bench command and one
# `--set full` capture of each dominant kernel.  Everything lands in gpurun_out/; summaries are copied to profiles/ afterwards.
set -u
mkdir -p gpurun_out
TAG=${1:-r2_33}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-extra"
KB="python scripts/kbench.py --ops fwd,bwd,apply --iters 2"
MC="python scripts/mlp_check.py --time"
timeout 600 $CMD > gpurun_out/${TAG}_plain_bench.log 2>&1 && timeout 300 $KB > gpurun_out/${TAG}_plain_kbench.log 2>&1 && timeout 300 $MC > gpurun_out/${TAG}_plain_mlp.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain_*.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain_kbench.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list exit: $?"
for K in dot_interaction_fwd_kernel dot_interaction_bwd_kernel seg_reduce_tiles_bulk_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_$K $KB > gpurun_out/${TAG}_ncu_$K.log 2>&1
  echo "ncu $K exit: $?"
done
# the 800 -> 512 forward product of the top tower: first dense_gemm launch of the timing pass
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 40 -c 1 -f -o gpurun_out/prof_${TAG}_dense_gemm_kernel $MC > gpurun_out/${TAG}_ncu_dense.log 2>&1
echo "ncu dense exit: $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:feature_bwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_din_feature_bwd python scripts/din_timeline.py --batch 16384 > gpurun_out/${TAG}_ncu_din.log 2>&1
echo "ncu din exit: $?"
ls -la gpurun_out/*${TAG}*.ncu-rep
