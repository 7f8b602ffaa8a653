#!/bin/bash
# ncu --set full over the 13 conv GEMM launches of one step (after 3 warm-up steps); raw CSV only.
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1 --in-flight 1"   # one lane: launch order = layer order
$CMD > $OUT/plain_conv_$TAG.log 2>&1 && \
ncu --set full --clock-control none -k regex:modconv_tc -s 42 -c 14 -f -o /tmp/prof_conv_$TAG $CMD > $OUT/ncu_conv_$TAG.log 2>&1
echo "ncu conv rc=$?"
ncu -i /tmp/prof_conv_$TAG.ncu-rep --page raw --csv > $OUT/prof_conv_$TAG.csv 2>/dev/null
ls -la $OUT/prof_conv_$TAG.csv
