#!/bin/bash
# r2_60: per-table id check tests + where the e2e loop's distance to `value` comes from
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "per_table_id_check or validate_ids or out_of_range" > gpurun_out/r2_60_pytest.log 2>&1
tail -3 gpurun_out/r2_60_pytest.log
RB_E2E_PROBE=1 timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2_60_bench.json 2> gpurun_out/r2_60_bench.err
tail -c 600 gpurun_out/r2_60_bench.err
python - <<'P'
import json
for l in open('gpurun_out/r2_60_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'])
P
