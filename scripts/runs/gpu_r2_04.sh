#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -m gpu -q --timeout 300 > gpurun_out/r2_04_pytest_mlp.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r2_04_pytest_mlp.log
timeout 120 python scripts/mlp_check.py --profile > gpurun_out/r2_04_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_gemm -s 6 -c 3 -f -o gpurun_out/r2_04_prof_gemm python scripts/mlp_check.py --profile > gpurun_out/r2_04_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2_04_ncu.log
