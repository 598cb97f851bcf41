#!/bin/bash
# r2_71: what the driver runs at round end — pytest -m gpu, smoke, default bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_71_pytest_gpu.log 2>&1
tail -2 gpurun_out/r2_71_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_71_smoke.log 2>&1
tail -2 gpurun_out/r2_71_smoke.log
timeout 600 python bench.py > gpurun_out/r2_71_bench.json 2> gpurun_out/r2_71_bench.err
tail -c 300 gpurun_out/r2_71_bench.err
python - <<'P'
import json
for l in open('gpurun_out/r2_71_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_alone'], d['roofline']['alone_ms'], d['sustained']['value'], d['clocks'], d['gpu_launches'])
P
