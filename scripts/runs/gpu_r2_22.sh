#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_22
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -x -k "fused_singleton" > gpurun_out/${T}_pytest_fused.log 2>&1; echo "pytest fused exit $?"; tail -15 gpurun_out/${T}_pytest_fused.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_mlp.py tests/test_gpu_train.py -m gpu -q --timeout 600 -x > gpurun_out/${T}_pytest_models.log 2>&1; echo "pytest models exit $?"; tail -5 gpurun_out/${T}_pytest_models.log
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err; head -c 230 gpurun_out/${T}_bench.json; echo
timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline.json > gpurun_out/${T}_timeline.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline.txt
RB_FUSED_UPDATE=0 timeout 900 python bench.py --no-cpu-baseline --no-extra --no-e2e > gpurun_out/${T}_bench_unfused.json 2> gpurun_out/${T}_bench_unfused.err
echo "bench unfused exit $?"; head -c 230 gpurun_out/${T}_bench_unfused.json; echo
timeout 600 python bench.py --no-cpu-baseline --no-extra --no-e2e --dist zipf > gpurun_out/${T}_bench_zipf.json 2> gpurun_out/${T}_bench_zipf.err
echo "bench zipf exit $?"; head -c 230 gpurun_out/${T}_bench_zipf.json; echo
