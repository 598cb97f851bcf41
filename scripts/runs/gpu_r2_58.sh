#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_58
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${T}_smoke.log
