#!/bin/bash
# ncu evidence for the bench step (run under gpurun, one GPU).  Numbers printed by bench.py under
# ncu are never bench values; only the per-launch durations / counters are used.
#   bash profiles/capture.sh <round-tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
# 1) every launch with its device time (cold-cache, serialised: compare SHARES of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
# 2) full counter set for the kernels of one step (the eager instrumented pass runs first; skip its
#    first step) ...
STEP="att_logits|edge_dot|run_apply|segment_softmax|gather_weights|layer_fwd|layer_bwd|csc_gather|input_bwd|bpr_|reduce_partials"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"$STEP" -s 24 -c 24 \
    -o $OUT/prof_$TAG -f $CMD --no-eval > $OUT/ncu_full_$TAG.log 2>&1
# 3) ... and for the evaluation kernels (operand split, scoring + top-K, list merge)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"score_topk|topk_merge|eval_split|eval_absmax" -c 6 \
    -o $OUT/prof_${TAG}_eval -f $CMD > $OUT/ncu_full_${TAG}_eval.log 2>&1
# 4) the TransR iteration (KG phase)
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"transr_|adam_" -s 8 -c 6 \
    -o $OUT/prof_${TAG}_kg -f python profiles/kg_once.py > $OUT/ncu_full_${TAG}_kg.log 2>&1
ls -la $OUT
