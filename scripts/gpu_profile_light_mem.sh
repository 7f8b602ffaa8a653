#!/bin/bash
# light ncu pass over the memory-bound launches (labelling, ToRGB, blur, mapping) of one step of config 2
TAG=${1:-r03}
OUT=gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
CMD="python bench.py --config 2 --extra-configs= --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1 --in-flight 1"
ncu --metrics $M --clock-control none -k regex:"label_|torgb|linear_|pixel_norm|assemble|modulat|demod|prescale|nchw_to" -s 90 -c 30 --csv --page raw --log-file $OUT/prof_light_mem_$TAG.csv $CMD > $OUT/ncu_light_mem_$TAG.log 2>&1
echo "ncu light mem rc=$?"
