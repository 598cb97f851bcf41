#!/bin/bash
set -u
mkdir -p gpurun_out
for pair in 0 1; do
  RB_DENSE_PAIR=$pair timeout 200 python scripts/mlp_check.py --time > gpurun_out/r2_10_mlp_check_pair$pair.log 2>&1; echo "check pair=$pair exit $?"
  grep -E "BAD|ALL|FAIL|^top1|^top2|^bot|Error|error|timed out" gpurun_out/r2_10_mlp_check_pair$pair.log | head -8
  RB_DENSE_PAIR=$pair timeout 200 python scripts/mlp_check.py --stats > gpurun_out/r2_10_stats_pair$pair.log 2>&1; echo "stats pair=$pair exit $?"; cat gpurun_out/r2_10_stats_pair$pair.log | tail -3
done
timeout 600 python -m pytest tests/test_gpu_mlp.py -m gpu -q --timeout 300 > gpurun_out/r2_10_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r2_10_pytest.log
for pair in 0 1; do
RB_DENSE_PAIR=$pair timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2_10_bench_pair$pair.json 2> gpurun_out/r2_10_bench_pair$pair.err
echo "bench pair=$pair exit $?"; tail -2 gpurun_out/r2_10_bench_pair$pair.err; head -c 230 gpurun_out/r2_10_bench_pair$pair.json; echo
done
# hot rows: local row copies through L1 (tuning build) vs the default, Zipf ids on one shared table
for lib in default localca; do
  if [ $lib = localca ]; then export RB_LIB_PATH=$PWD/recommender_b200/lib/librecsys_b200_localca.so; fi
  timeout 300 python scripts/kbench.py --ops fwd,bwd --tables 1 --dist zipf --tag $lib > gpurun_out/r2_10_kbench_zipf_t1_$lib.json 2>&1
  timeout 300 python scripts/kbench.py --ops fwd,bwd --tables 26 --dist uniform --tag $lib > gpurun_out/r2_10_kbench_uniform_t26_$lib.json 2>&1
  tail -1 gpurun_out/r2_10_kbench_zipf_t1_$lib.json; tail -1 gpurun_out/r2_10_kbench_uniform_t26_$lib.json
done
