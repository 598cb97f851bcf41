#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_34
timeout 300 python scripts/kbench.py --ops apply --tag flat_ring4 > gpurun_out/${T}_kbench_ring4.json 2> gpurun_out/${T}_kbench.err; cat gpurun_out/${T}_kbench_ring4.json
for v in ring3 ring6; do
RB_LIB_PATH=recommender_b200/lib/librecsys_b200_$v.so timeout 300 python scripts/kbench.py --ops apply --tag flat_$v > gpurun_out/${T}_kbench_$v.json 2>> gpurun_out/${T}_kbench.err; cat gpurun_out/${T}_kbench_$v.json
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "sparse or dedup or chain or fused or consumer or uses" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; head -c 230 gpurun_out/${T}_bench.json; echo
