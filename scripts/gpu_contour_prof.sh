#!/bin/bash
# contour stage alone: event-timed, then an ncu launch list of the stage on synthetic masks and inside the dataset leg
TAG=${1:-r02n}
timeout 300 python scripts/contour_stage_bench.py > gpurun_out/${TAG}_contour_stage.jsonl 2> gpurun_out/${TAG}_contour_stage.err; echo rc=$?; cat gpurun_out/${TAG}_contour_stage.jsonl; tail -3 gpurun_out/${TAG}_contour_stage.err
M=gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none -k regex:"ct_" --csv --page raw --log-file gpurun_out/${TAG}_contour_ncu.csv python scripts/contour_stage_bench.py --iters 1 > gpurun_out/${TAG}_contour_ncu.log 2>&1; echo ncu rc=$?
timeout 900 ncu --metrics $M --clock-control none -k regex:"ct_" -c 600 --csv --page raw --log-file gpurun_out/${TAG}_contour_leg_ncu.csv python bench.py --leg dataset --contours device --steps 2 > gpurun_out/${TAG}_contour_leg_ncu.log 2>&1; echo ncu leg rc=$?
