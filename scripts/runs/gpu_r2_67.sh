#!/bin/bash
# r2_67: Dense kernels launched programmatically (prologue under the previous kernel's tail) — parity, then the step with RB_PDL=1/0
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_67_pytest.log 2>&1
tail -3 gpurun_out/r2_67_pytest.log
for pdl in 1 0 1 0; do
RB_PDL=$pdl timeout 300 python bench.py --no-cpu-baseline --no-extra --no-e2e --sustain-seconds 0 > gpurun_out/r2_67_bench_pdl$pdl.json 2> gpurun_out/r2_67_bench.err
tail -c 300 gpurun_out/r2_67_bench.err
python - gpurun_out/r2_67_bench_pdl$pdl.json $pdl <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('pdl',sys.argv[2], d['value'], d['ms_per_step'], d['eager_ms_per_step'])
P
done
