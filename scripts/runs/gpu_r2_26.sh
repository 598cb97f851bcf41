#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_26
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q --timeout 600 -k "fused" > gpurun_out/${T}_pytest_fused.log 2>&1; echo "pytest fused exit $?"; tail -8 gpurun_out/${T}_pytest_fused.log
