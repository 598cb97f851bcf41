#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_17_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/r2_17_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_17_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2_17_smoke.log
timeout 900 python bench.py > gpurun_out/r2_17_bench.json 2> gpurun_out/r2_17_bench.err
echo "bench exit $?"; tail -2 gpurun_out/r2_17_bench.err; head -c 230 gpurun_out/r2_17_bench.json; echo
