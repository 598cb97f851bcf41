#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_20
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_mlp.py -m gpu -q --timeout 600 -x > gpurun_out/${T}_pytest_models.log 2>&1; echo "pytest models exit $?"; tail -3 gpurun_out/${T}_pytest_models.log
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err; head -c 230 gpurun_out/${T}_bench.json; echo
timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline.json > gpurun_out/${T}_timeline.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline.txt
RB_PRESORT_AT=after_lookup timeout 900 python bench.py --no-cpu-baseline --no-extra --no-e2e > gpurun_out/${T}_bench_late.json 2> gpurun_out/${T}_bench_late.err
echo "bench late exit $?"; head -c 230 gpurun_out/${T}_bench_late.json; echo
RB_PRESORT_AT=after_lookup timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline_late.json > gpurun_out/${T}_timeline_late.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline_late.txt
RB_PRIO_MAIN=-3 RB_PRIO_WGRAD=-1 RB_PRIO_SIDE=-2 timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline_p312.json > gpurun_out/${T}_timeline_p312.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline_p312.txt
