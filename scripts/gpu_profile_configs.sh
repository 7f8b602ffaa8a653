#!/bin/bash
# ncu --set full of the conv GEMM + blur launches of one step of BASELINE configs 3 and 4 (one lane: launch order = layer order).
# Usage: scripts/gpu_profile_configs.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for CFG in 3 4; do
  CMD="python bench.py --config $CFG --extra-configs= --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1 --in-flight 1"
  # kernels matched per step: 512^2: 15 StyledConv (+1 strip launch of the Cout=128 up-conv) + 7 blur = 23; 1024^2: 17 (+1) + 8 = 26
  if [ $CFG = 3 ]; then PER=23; else PER=26; fi
  $CMD > $OUT/plain_cfg${CFG}_$TAG.log 2>&1 && \
  ncu --set full --clock-control none -k regex:"modconv_tc|blur_act_split" -s $((3 * PER)) -c $PER -f -o /tmp/prof_cfg${CFG}_$TAG $CMD > $OUT/ncu_cfg${CFG}_$TAG.log 2>&1
  echo "ncu cfg$CFG rc=$?"
  ncu -i /tmp/prof_cfg${CFG}_$TAG.ncu-rep --page raw --csv > $OUT/prof_cfg${CFG}_$TAG.csv 2>/dev/null
done
ls -la $OUT | tail -8
