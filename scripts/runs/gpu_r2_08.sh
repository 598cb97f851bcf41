#!/bin/bash
set -u
mkdir -p gpurun_out
RB_DENSE_PAIR=1 timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_08_mlp_check_pair.log 2>&1; echo "pair exit $?"
grep -E "BAD|ALL|FAIL|^top|^bot|Error|error|timed out" gpurun_out/r2_08_mlp_check_pair.log | head -20
RB_DENSE_PAIR=0 timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_08_mlp_check_single.log 2>&1; echo "single exit $?"
grep -E "BAD|ALL|FAIL|^top|^bot|Error|error|timed out" gpurun_out/r2_08_mlp_check_single.log | head -20
