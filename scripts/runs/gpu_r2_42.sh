#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_replicated.py tests/test_gpu_din.py -m gpu -q --timeout 600 > gpurun_out/r2_42_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2_42_pytest.log
