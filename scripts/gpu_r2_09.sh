#!/bin/bash
set -u
mkdir -p gpurun_out
for pair in 0 1; do
  RB_DENSE_PAIR=$pair timeout 200 python scripts/mlp_check.py --stats > gpurun_out/r2_09_stats_pair$pair.log 2>&1; echo "stats pair=$pair exit $?"; cat gpurun_out/r2_09_stats_pair$pair.log | tail -4
  RB_DENSE_PAIR=$pair timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_09_mlp_check_pair$pair.log 2>&1; echo "check pair=$pair exit $?"
  grep -E "BAD|ALL|FAIL|^top1|^top2|Error|error|timed out" gpurun_out/r2_09_mlp_check_pair$pair.log | head -8
done
