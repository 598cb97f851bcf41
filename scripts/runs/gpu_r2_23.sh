#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_23
timeout 300 python scripts/kbench.py --ops bwd_update --iters 10 > gpurun_out/${T}_kbench.json 2> gpurun_out/${T}_kbench.err; echo "kbench exit $?"; cat gpurun_out/${T}_kbench.json; tail -3 gpurun_out/${T}_kbench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dot_interaction_bwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_${T}_bwd_update python scripts/kbench.py --ops bwd_update --iters 2 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*${T}*
