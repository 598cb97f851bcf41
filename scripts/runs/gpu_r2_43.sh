#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_43
timeout 300 python scripts/kbench.py --ops apply --tag evict_first > gpurun_out/${T}_kbench_evict.json 2> gpurun_out/${T}_kbench.err; cat gpurun_out/${T}_kbench_evict.json
RB_LIB_PATH=recommender_b200/lib/librecsys_b200_noevict.so timeout 300 python scripts/kbench.py --ops apply --tag noevict > gpurun_out/${T}_kbench_noevict.json 2>> gpurun_out/${T}_kbench.err; cat gpurun_out/${T}_kbench_noevict.json
timeout 300 python scripts/kbench.py --ops apply --tag evict_first_zipf --dist zipf > gpurun_out/${T}_kbench_evict_zipf.json 2>> gpurun_out/${T}_kbench.err; cat gpurun_out/${T}_kbench_evict_zipf.json
RB_LIB_PATH=recommender_b200/lib/librecsys_b200_noevict.so timeout 300 python scripts/kbench.py --ops apply --tag noevict_zipf --dist zipf > gpurun_out/${T}_kbench_noevict_zipf.json 2>> gpurun_out/${T}_kbench.err; cat gpurun_out/${T}_kbench_noevict_zipf.json
tail -3 gpurun_out/${T}_kbench.err
