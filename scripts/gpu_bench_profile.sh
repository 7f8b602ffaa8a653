#!/bin/bash
# Run on the B200 box through gpurun: smoke, bench (N=1), then ncu launch list + full captures of the hot kernels.
# Usage: scripts/gpu_bench_profile.sh <round-tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench_$TAG.json
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1"
$CMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:modconv_tc_kernel -s 47 -c 5 -f -o $OUT/prof_conv_$TAG $CMD > $OUT/ncu_conv_$TAG.log 2>&1
echo "ncu conv rc=$?"
$CMD > $OUT/plain3_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"label_native|blur_act_split|torgb_kernel" -s 51 -c 17 -f -o $OUT/prof_mem_$TAG $CMD > $OUT/ncu_mem_$TAG.log 2>&1
echo "ncu mem rc=$?"
ls -la $OUT
