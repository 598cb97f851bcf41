#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_37
S=$(date +%s); timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $? wall $(( $(date +%s) - S )) s"; head -c 230 gpurun_out/${T}_bench.json; echo
S=$(date +%s); timeout 1200 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
echo "bench ref exit $? wall $(( $(date +%s) - S )) s"; head -c 600 gpurun_out/${T}_bench_reference.json; echo
