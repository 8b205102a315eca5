#!/bin/bash
# One 1-GPU box session: GPU tests, the bench lines of every single-GPU config, the reference arm, the ncu captures.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_session.sh r2'
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt
rm -f $OUT/parity_fullsize_$TAG.jsonl
( time KGAT_PARITY_LOG=$OUT/parity_fullsize_$TAG.jsonl timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
for c in last-fm yelp2018 amazon-book-10x; do
  timeout 600 python bench.py --config $c --no-cpu-baseline > $OUT/bench_${TAG}_$c.json 2> $OUT/bench_${TAG}_$c.err
done
bash profiles/capture.sh $TAG > $OUT/capture_$TAG.log 2>&1
ls -la $OUT | tail -30
