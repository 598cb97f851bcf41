#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q --timeout 300 > gpurun_out/r2_07_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r2_07_pytest.log
timeout 300 python scripts/timeline.py --out gpurun_out/r2_07_timeline.json > gpurun_out/r2_07_timeline.txt 2>&1; echo "timeline exit $?"; tail -2 gpurun_out/r2_07_timeline.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_07_bench.json 2> gpurun_out/r2_07_bench.err
echo "bench exit $?"; tail -3 gpurun_out/r2_07_bench.err; head -c 250 gpurun_out/r2_07_bench.json; echo
RB_SEG_BULK=0 timeout 300 python scripts/kbench.py --ops update > gpurun_out/r2_07_kbench_ring.json 2>&1; RB_SEG_BULK=1 timeout 300 python scripts/kbench.py --ops update > gpurun_out/r2_07_kbench_bulk.json 2>&1
tail -1 gpurun_out/r2_07_kbench_ring.json; tail -1 gpurun_out/r2_07_kbench_bulk.json
