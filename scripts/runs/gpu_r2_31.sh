#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_51
timeout 900 python -m pytest tests/test_gpu_din.py tests/test_gpu_mlp.py -m gpu -q --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/${T}_pytest.log
timeout 600 python scripts/din_timeline.py > gpurun_out/${T}_din_timeline.txt 2>&1; echo "exit $?"; head -22 gpurun_out/${T}_din_timeline.txt | cut -c1-140
