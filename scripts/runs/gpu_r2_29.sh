#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_29
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err; head -c 230 gpurun_out/${T}_bench.json; echo
python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['value'], 'sustained', d.get('value_sustained'), 'cpu', d.get('cpu_baseline'))
print(json.dumps(d['extra'].get('din_cfg4_attention')))
"
