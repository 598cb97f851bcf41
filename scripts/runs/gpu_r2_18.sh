#!/bin/bash
set -u
mkdir -p gpurun_out
for rs in 0 1; do
  RB_DENSE_RESIDENT=$rs timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_18_mlp_check_res$rs.log 2>&1; echo "check resident=$rs exit $?"
  grep -E "BAD|ALL|FAIL|^top1|^top2|^bot|Error|error|timed out" gpurun_out/r2_18_mlp_check_res$rs.log | head -8
done
RB_DENSE_RESIDENT=2 timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "dense or mlp or benchmarked or deepfm or stay_inside" > gpurun_out/r2_18_pytest_forced.log 2>&1; echo "pytest (forced resident) exit $?"; tail -4 gpurun_out/r2_18_pytest_forced.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_18_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r2_18_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2_18_bench.json 2> gpurun_out/r2_18_bench.err
echo "bench exit $?"; tail -2 gpurun_out/r2_18_bench.err; head -c 230 gpurun_out/r2_18_bench.json; echo
timeout 300 python scripts/timeline.py --out gpurun_out/r2_18_timeline.json > gpurun_out/r2_18_timeline.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/r2_18_timeline.txt
