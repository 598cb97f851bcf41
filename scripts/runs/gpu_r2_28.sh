#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_28
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_din.py -m gpu -q --timeout 600 -x -k "forward_backward_golden" > gpurun_out/${T}_pytest_din.log 2>&1; echo "pytest din exit $?"; grep -E "RecsysError|Error:|din.py:[0-9]+|ops.py:[0-9]+" gpurun_out/${T}_pytest_din.log | head -12
