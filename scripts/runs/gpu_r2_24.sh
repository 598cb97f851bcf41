#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_25
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err; head -c 230 gpurun_out/${T}_bench.json; echo
