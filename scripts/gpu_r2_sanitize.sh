#!/bin/bash
# usage: gpu_r2_sanitize.sh memcheck|racecheck   (ONE tool per gpurun call, B200_PROFILING.md)
set -u
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL='not full_size and not 65536 and not two_real_ranks and not graphed and not tfrecord'
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mlp.py -m gpu -q -x --timeout 600 -k "$SEL" > gpurun_out/r2_sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_sanitize_plain.log; exit 1; }
tail -2 gpurun_out/r2_sanitize_plain.log
timeout 3000 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/r2_sanitize_$TOOL.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mlp.py -m gpu -q -x --timeout 2400 -k "$SEL" > gpurun_out/r2_sanitize_${TOOL}_pytest.log 2>&1
echo "$TOOL exit $?"; tail -3 gpurun_out/r2_sanitize_${TOOL}_pytest.log; tail -5 gpurun_out/r2_sanitize_$TOOL.log
