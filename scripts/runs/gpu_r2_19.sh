#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_19
timeout 200 python scripts/mlp_check.py --time > gpurun_out/${T}_mlp_check.log 2>&1; echo "mlp_check exit $?"
grep -E "BAD|ALL|FAIL|^top1|^top2|^bot|Error|error|timed out" gpurun_out/${T}_mlp_check.log | head -8
for pad in 0 12 40; do
  RB_SEG_PAD_KB=$pad timeout 200 python scripts/kbench.py --ops apply --tag pad$pad > gpurun_out/${T}_kbench_apply_pad$pad.json 2> gpurun_out/${T}_kbench.err; echo "kbench pad=$pad exit $?"; cat gpurun_out/${T}_kbench_apply_pad$pad.json
done
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/${T}_bench.err; head -c 230 gpurun_out/${T}_bench.json; echo
timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline.json > gpurun_out/${T}_timeline.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline.txt
RB_PRIO_MAIN=-1 timeout 900 python bench.py --no-cpu-baseline --no-extra --no-e2e > gpurun_out/${T}_bench_prio.json 2> gpurun_out/${T}_bench_prio.err
echo "bench prio exit $?"; head -c 230 gpurun_out/${T}_bench_prio.json; echo
RB_PRIO_MAIN=-1 timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline_prio.json > gpurun_out/${T}_timeline_prio.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline_prio.txt
RB_PRIO_MAIN=-2 RB_PRIO_WGRAD=-1 timeout 300 python scripts/timeline.py --out gpurun_out/${T}_timeline_prio2.json > gpurun_out/${T}_timeline_prio2.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/${T}_timeline_prio2.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/${T}_pytest_gpu.log
