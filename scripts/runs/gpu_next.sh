#!/bin/bash
# First GPU call of the next round: the measurements DESIGN.md §6b lists as open.  Everything lands in gpurun_out/.
#   gpurun --timeout 600 -- 'bash scripts/gpu_next.sh'
set -u
mkdir -p gpurun_out
# 1. hot rows: the headline step with Zipf(1.05) ids + 2 % forced id 0 (SURVEY §8d asks for both distributions)
timeout 200 python bench.py --dist zipf --no-cpu-baseline > gpurun_out/bench_zipf.json 2> gpurun_out/bench_zipf.err
tail -c 600 gpurun_out/bench_zipf.json
# 2. the same with one shared table (the reference's layout, T = 1): chains of duplicates are ~26x longer
timeout 200 python bench.py --tables 1 --dist zipf --no-cpu-baseline --no-e2e > gpurun_out/bench_zipf_t1.json 2> gpurun_out/bench_zipf_t1.err
tail -c 600 gpurun_out/bench_zipf_t1.json
# 2b. opt-in collapsed towers (linear hidden layers => one affine map): same step, MLP GEMMs replaced by three passes
timeout 200 python bench.py --collapse-mlp --no-cpu-baseline > gpurun_out/bench_collapse_mlp.json 2> gpurun_out/bench_collapse_mlp.err
tail -c 600 gpurun_out/bench_collapse_mlp.json
# 3. why the trainer loop runs at 6.9 ms per step at B = 65536: launch list of one short run (shares, not absolutes)
timeout 120 python scripts/train_bench.py --lines 1000000 --only dlrm_b65536 > gpurun_out/train_b65536.json 2> gpurun_out/train_b65536.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train_b65536.csv \
    python scripts/train_bench.py --lines 400000 --only dlrm_b65536 > gpurun_out/ncu_train.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_train_b65536.csv > gpurun_out/launches_train_b65536.md 2>&1 || true
head -30 gpurun_out/launches_train_b65536.md
