#!/bin/bash
set -u
mkdir -p gpurun_out
for pair in 0 1; do
  RB_DENSE_PAIR=$pair timeout 200 python scripts/mlp_check.py --stats > gpurun_out/r2_09_stats_pair$pair.log 2>&1; echo "stats pair=$pair exit $?"; cat gpurun_out/r2_09_stats_pair$pair.log | tail -4
  RB_DENSE_PAIR=$pair timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_09_mlp_check_pair$pair.log 2>&1; echo "check pair=$pair exit $?"
  grep -E "BAD|ALL|FAIL|^top1|^top2|Error|error|timed out" gpurun_out/r2_09_mlp_check_pair$pair.log | head -8
done
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "mlp or dense or deepfm or benchmarked or overflow or peer_memory" > gpurun_out/r2_09_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r2_09_pytest.log
timeout 600 python bench.py > gpurun_out/r2_09_bench.json 2> gpurun_out/r2_09_bench.err
echo "bench exit $?"; tail -3 gpurun_out/r2_09_bench.err; head -c 300 gpurun_out/r2_09_bench.json; echo
