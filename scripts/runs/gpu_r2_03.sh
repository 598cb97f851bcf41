#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_03_mlp_check.log 2>&1; echo "mlp_check exit $?"
grep -E "BAD|ALL|FAIL|^top|^bot" gpurun_out/r2_03_mlp_check.log
timeout 600 python -m pytest tests/test_gpu_mlp.py -m gpu -q -x --timeout 300 > gpurun_out/r2_03_pytest_mlp.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r2_03_pytest_mlp.log
for be in tcgen05 cublas; do
  timeout 300 python bench.py --mlp-backend $be --no-cpu-baseline > gpurun_out/r2_03_bench_$be.json 2> gpurun_out/r2_03_bench_$be.err
  echo "bench $be exit $?"; tail -3 gpurun_out/r2_03_bench_$be.err; head -c 300 gpurun_out/r2_03_bench_$be.json; echo
done
