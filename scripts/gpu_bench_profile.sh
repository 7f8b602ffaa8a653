#!/bin/bash
# Run on the B200 box through gpurun: smoke, bench (N=1), then ncu launch list + full captures of the hot kernels.
# Usage: scripts/gpu_bench_profile.sh <round-tag> [profile-only]
# gpurun copies back at most 64 MiB: the .ncu-rep files are converted to raw CSV on the box and only one small
# report (a single launch of the dominant kernel, with source) is kept.
TAG=${1:-r01}
ONLY=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
if [ "$ONLY" = "all" ]; then
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -c 1500 $OUT/bench_$TAG.json
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"
fi
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1 --in-flight 1"   # one lane: launch order = layer order
# launches of my library per resident step: 12 mapping + 2 input + 14 conv + 6 blur + 6 torgb + 4 label = 44
$CMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
# all 14 conv GEMM launches of the 4th step (13 layers; the Cout = 128 up-conv issues two; 3 warm-up steps skipped)
$CMD > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none -k regex:modconv_tc -s 42 -c 14 -f -o /tmp/prof_conv_$TAG $CMD > $OUT/ncu_conv_$TAG.log 2>&1
echo "ncu conv rc=$?"
ncu -i /tmp/prof_conv_$TAG.ncu-rep --page raw --csv > $OUT/prof_conv_$TAG.csv 2>/dev/null
# the memory-bound kernels of the 4th step: 6 blur + 5 torgb + 4 label (the 64^2 and 256^2 ToRGBs are fused into label launches)
$CMD > $OUT/plain3_$TAG.log 2>&1 && \
ncu --set full --clock-control none -k regex:"label_|blur_act_split|torgb_" -s 45 -c 15 -f -o /tmp/prof_mem_$TAG $CMD > $OUT/ncu_mem_$TAG.log 2>&1
echo "ncu mem rc=$?"
ncu -i /tmp/prof_mem_$TAG.ncu-rep --page raw --csv > $OUT/prof_mem_$TAG.csv 2>/dev/null
# one launch of the dominant kernel (last plain conv, 128->128 at 256^2) with source, kept as a report
$CMD > $OUT/plain4_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:modconv_tc -s 55 -c 1 -f -o $OUT/prof_conv_last_$TAG $CMD > $OUT/ncu_conv_last_$TAG.log 2>&1
echo "ncu conv-last rc=$?"
du -sh $OUT; ls -la $OUT | head -40
ncu -i $OUT/prof_conv_last_$TAG.ncu-rep --page source --csv > $OUT/prof_conv_last_source_$TAG.csv 2>/dev/null
