#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_38
timeout 900 python bench.py --no-cpu-baseline --no-extra --sustain-seconds 0 --steps 100 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
"
