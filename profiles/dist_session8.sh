#!/bin/bash
# 8-GPU box session: the multi-GPU parity tests (world 2, 4, 8), then bench.py at 8 ranks (default exchange).
set -u
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 500 python -m pytest tests/test_gpu_dist.py -x -q -m gpu ) > $OUT/pytest_dist_${TAG}_n8.log 2>&1
tail -4 $OUT/pytest_dist_${TAG}_n8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_${TAG}_n8.json 2> $OUT/bench_${TAG}_n8.err
echo "rc=$?"; tail -c 1500 $OUT/bench_${TAG}_n8.json
