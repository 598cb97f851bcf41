#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_53
timeout 900 python -m pytest tests/test_gpu_din.py -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_din.log 2>&1; echo "pytest din exit $?"; tail -30 gpurun_out/${T}_pytest_din.log
