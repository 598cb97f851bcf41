#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/timeline.py --out gpurun_out/r2_46_timeline.json > gpurun_out/r2_46_timeline.txt 2>&1; echo "timeline exit $?"; tail -1 gpurun_out/r2_46_timeline.txt
