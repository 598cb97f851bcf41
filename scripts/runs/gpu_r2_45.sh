#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_47
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_models.py -m gpu -q --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err; head -c 230 gpurun_out/${T}_bench.json; echo
RB_FUSED_HEAD=0 timeout 900 python bench.py --no-cpu-baseline --no-extra --no-e2e --sustain-seconds 0 > gpurun_out/${T}_bench_unfused_head.json 2> gpurun_out/${T}_bench_unfused_head.err
echo "bench (unfused head) exit $?"; head -c 230 gpurun_out/${T}_bench_unfused_head.json; echo
timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline.json > gpurun_out/${T}_timeline.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline.txt
