#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_49
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${T}_smoke.log
S=$(date +%s); timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $? wall $(( $(date +%s) - S )) s"; tail -2 gpurun_out/${T}_bench.err; head -c 230 gpurun_out/${T}_bench.json; echo
python -c "
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench.json') if l.startswith('{')][-1])
print('e2e', d['e2e']['value'], 'sustained', d.get('value_sustained'))
print(json.dumps(d['extra'].get('din_cfg4_attention'))[:1800])
"
