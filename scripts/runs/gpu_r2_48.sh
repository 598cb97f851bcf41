#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_48
for P in 1 0 -2 -3; do
RB_PRIO_SORT=$P timeout 600 python bench.py --no-cpu-baseline --no-extra --no-e2e --sustain-seconds 0 --steps 60 > gpurun_out/${T}_bench_sort$P.json 2> gpurun_out/${T}_bench_sort$P.err
echo "sort prio $P exit $?: $(python -c "import json; d=json.loads([l for l in open('gpurun_out/${T}_bench_sort$P.json') if l.startswith('{')][-1]); print(d['ms_per_step'])")"
done
RB_PRIO_SORT=-3 timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline_sort-3.json > gpurun_out/${T}_timeline_sort-3.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline_sort-3.txt
