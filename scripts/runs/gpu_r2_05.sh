#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_05_mlp_check.log 2>&1; echo "mlp_check exit $?"
grep -E "BAD|ALL|FAIL|^top|^bot" gpurun_out/r2_05_mlp_check.log
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_models.py -m gpu -q --timeout 300 > gpurun_out/r2_05_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r2_05_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_05_bench.json 2> gpurun_out/r2_05_bench.err
echo "bench exit $?"; tail -3 gpurun_out/r2_05_bench.err; head -c 300 gpurun_out/r2_05_bench.json; echo
