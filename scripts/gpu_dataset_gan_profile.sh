#!/bin/bash
# ncu --set full over the kernels of one DatasetGAN labelling call (B=8 to keep the replay short); raw CSV only.
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --leg dataset_gan --batch 8"
$CMD > $OUT/plain_dgan.log 2>&1 && \
ncu --set full --clock-control none -k regex:"dataset_gan_tail|nchw_to_nhwc_split|modconv_tc_kernel" -s 70 -c 22 -f -o /tmp/prof_dgan $CMD > $OUT/ncu_dgan.log 2>&1
echo "ncu dgan rc=$?"
ncu -i /tmp/prof_dgan.ncu-rep --page raw --csv > $OUT/prof_dgan_r01f.csv 2>/dev/null
ls -la $OUT/prof_dgan_r01f.csv
