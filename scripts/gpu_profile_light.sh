#!/bin/bash
# light ncu pass (a handful of metrics, few replays) over the conv + blur launches of one step of the given configs
# Usage: scripts/gpu_profile_light.sh <tag> "<configs>"
TAG=${1:-r02}
CFGS=${2:-"3 4"}
OUT=gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
for CFG in $CFGS; do
  CMD="python bench.py --config $CFG --extra-configs= --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1 --in-flight 1"
  if [ $CFG = 2 ]; then PER=20; elif [ $CFG = 3 ]; then PER=22; else PER=25; fi
  ncu --metrics $M --clock-control none -k regex:"modconv_tc|blur_act_split" -s $((3 * PER)) -c $PER --csv --page raw --log-file $OUT/prof_light_cfg${CFG}_$TAG.csv $CMD > $OUT/ncu_light_cfg${CFG}_$TAG.log 2>&1
  echo "ncu light cfg$CFG rc=$?"
done
