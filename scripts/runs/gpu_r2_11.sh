#!/bin/bash
set -u
mkdir -p gpurun_out
for pair in 0 1; do
  RB_DENSE_PAIR=$pair timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_11_mlp_check_pair$pair.log 2>&1; echo "check pair=$pair exit $?"
  grep -E "BAD|ALL|FAIL|^top1|^top2|^bot|Error|error|timed out" gpurun_out/r2_11_mlp_check_pair$pair.log | head -8
  RB_DENSE_PAIR=$pair timeout 200 python scripts/mlp_check.py --stats > gpurun_out/r2_11_stats_pair$pair.log 2>&1; echo "stats pair=$pair exit $?"; cat gpurun_out/r2_11_stats_pair$pair.log | tail -3
done
for t in 26; do for rc in l2 l1; do
  timeout 300 python scripts/kbench.py --ops fwd,bwd --tables $t --dist zipf --row-cache $rc > gpurun_out/r2_11_kbench_zipf_t${t}_$rc.json 2>&1; tail -1 gpurun_out/r2_11_kbench_zipf_t${t}_$rc.json
done; done
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_11_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r2_11_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_11_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2_11_smoke.log
RB_DENSE_PAIR=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_11_bench.json 2> gpurun_out/r2_11_bench.err
echo "bench exit $?"; tail -2 gpurun_out/r2_11_bench.err; head -c 230 gpurun_out/r2_11_bench.json; echo
