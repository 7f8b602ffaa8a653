#!/bin/bash
# light ncu pass (time, DRAM / L2 / L1 throughput, occupancy, registers) over the launches of ONE sis_contour_stage call on
# the dataset leg's masks
TAG=${1:-r04}
OUT=gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:"ct_" -s 150 -c 75 --csv --page raw --log-file $OUT/prof_light_contours_$TAG.csv python bench.py --leg dataset --contours device --steps 2 > $OUT/ncu_light_contours_$TAG.log 2>&1
echo "ncu light contours rc=$?"
