#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, then (only if the plain run exited 0)
# the ncu launch list of the same bench command.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit: $?"
tail -c 6000 gpurun_out/bench.json; tail -20 gpurun_out/bench.err
if [ "${1:-}" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
  timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
  echo "ncu exit: $?"
  tail -3 gpurun_out/ncu.log
fi
