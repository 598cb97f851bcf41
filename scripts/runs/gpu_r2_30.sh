#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/din_timeline.py > gpurun_out/r2_30_din_timeline.txt 2>&1; echo "exit $?"; head -40 gpurun_out/r2_30_din_timeline.txt
