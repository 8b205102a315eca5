#!/bin/bash
# ncu evidence for the bench step (run under gpurun, one GPU).  Numbers printed by bench.py under
# ncu are never bench values; only the per-launch durations / counters are used.
#   bash profiles/capture.sh <round-tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
# 1) every launch with its device time (cold-cache, serialised: compare SHARES of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
# 2) full counter set for our kernels, second step onwards
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"att_logits|layer_fwd|layer_bwd|input_bwd|segment_softmax|score_topk|transr_" -s 10 -c 11 \
    -o $OUT/prof_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
