#!/bin/bash
# Round-2 first call: the distributions SURVEY §8d asks for, on the round-1 code (baseline for this round's changes).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for cfg in "zipf:--dist zipf" "zipf_t1:--tables 1 --dist zipf" "uniform_t1:--tables 1" "collapse:--collapse-mlp"; do
  tag=${cfg%%:*}; flags=${cfg#*:}
  timeout 240 python bench.py $flags --no-cpu-baseline > gpurun_out/r2_00_bench_$tag.json 2> gpurun_out/r2_00_bench_$tag.err
  echo "$tag exit $?"; tail -c 400 gpurun_out/r2_00_bench_$tag.json | head -c 400; echo
done
