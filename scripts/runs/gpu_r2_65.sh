#!/bin/bash
# r2_65: border kernels of the row update launched programmatically (griddepcontrol) — parity, then A/B alone and in the step
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -x > gpurun_out/r2_65_pytest.log 2>&1
tail -3 gpurun_out/r2_65_pytest.log
for pdl in 1 0 1 0; do
  for dist in uniform zipf; do
    RB_SEG_PDL=$pdl timeout 200 python scripts/kbench.py --ops apply --iters 30 --dist $dist --tag pdl${pdl}_$dist 2>&1 | tail -1
  done
done > gpurun_out/r2_65_kbench.txt 2>&1
cat gpurun_out/r2_65_kbench.txt
for pdl in 1 0; do
RB_SEG_PDL=$pdl timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2_65_bench_pdl$pdl.json 2> gpurun_out/r2_65_bench.err
tail -c 300 gpurun_out/r2_65_bench.err
python - gpurun_out/r2_65_bench_pdl$pdl.json <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['alone_ms'], d['roofline']['frac_alone'], d['sustained']['value'])
P
done
