#!/bin/bash
set -u
N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q --timeout 600 -k "bucket or peer_memory or sharded or hash or two_real_ranks" > gpurun_out/r2_15_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2_15_pytest.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 --config2-sharded --sustain-seconds 0 --no-e2e > gpurun_out/r2_15_n2_cfg2.json 2> gpurun_out/r2_15_n2_cfg2.err
echo "cfg2 exit $?"; grep '^{' gpurun_out/r2_15_n2_cfg2.json | head -c 220; echo
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 --dist zipf --sustain-seconds 0 --no-e2e > gpurun_out/r2_15_n2_cfg3_zipf.json 2> gpurun_out/r2_15_n2_cfg3_zipf.err
echo "cfg3 zipf exit $?"; tail -2 gpurun_out/r2_15_n2_cfg3_zipf.err; grep '^{' gpurun_out/r2_15_n2_cfg3_zipf.json | head -c 220; echo
