#!/bin/bash
# r2_72: small kernels release their dependents right after their own wait (default build) vs at their end (variant) — parity + A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_models.py tests/test_gpu_train.py -q -m gpu -x > gpurun_out/r2_72_pytest.log 2>&1
tail -2 gpurun_out/r2_72_pytest.log
for v in default norelease default norelease; do
  if [ $v = default ]; then unset RB_LIB_PATH; else export RB_LIB_PATH=$PWD/recommender_b200/lib/librecsys_b200_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-extra --no-e2e --sustain-seconds 0 > gpurun_out/r2_72_bench_$v.json 2> gpurun_out/r2_72_bench.err
  tail -c 200 gpurun_out/r2_72_bench.err
  python - gpurun_out/r2_72_bench_$v.json $v <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('variant',sys.argv[2], d['value'], d['ms_per_step'])
P
done
