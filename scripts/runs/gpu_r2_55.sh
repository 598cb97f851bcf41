#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_55
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -k "graph" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench.json') if l.startswith('{')][-1])
print('bound:', d['ms_per_step'], d['value']/1e6, 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], 'sustained', d['value_sustained']/1e6)"
timeout 900 python bench.py --no-cpu-baseline --no-extra --no-bind-inputs > gpurun_out/${T}_bench_copy.json 2> gpurun_out/${T}_bench_copy.err
echo "bench copy exit $?"; python -c "
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_copy.json') if l.startswith('{')][-1])
print('copy :', d['ms_per_step'], d['value']/1e6, 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], 'sustained', d['value_sustained']/1e6)"
