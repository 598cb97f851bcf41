#!/bin/bash
set -u
mkdir -p gpurun_out
for pf in 0 1; do
  RB_DENSE_PREFETCH=$pf timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_12_mlp_check_pf$pf.log 2>&1; echo "check prefetch=$pf exit $?"
  grep -E "BAD|ALL|FAIL|^top1|^top2|^bot|Error|error|timed out" gpurun_out/r2_12_mlp_check_pf$pf.log | head -8
  RB_DENSE_PREFETCH=$pf timeout 200 python scripts/mlp_check.py --stats > gpurun_out/r2_12_stats_pf$pf.log 2>&1; echo "stats prefetch=$pf exit $?"; cat gpurun_out/r2_12_stats_pf$pf.log | tail -3
done
timeout 600 python -m pytest tests/test_gpu_mlp.py -m gpu -q --timeout 300 > gpurun_out/r2_12_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_12_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_12_bench.json 2> gpurun_out/r2_12_bench.err
echo "bench exit $?"; tail -2 gpurun_out/r2_12_bench.err; head -c 230 gpurun_out/r2_12_bench.json; echo
