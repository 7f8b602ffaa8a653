#!/bin/bash
# BASELINE config 5: kernel sweeps with CUDA-event timings, then the same sweeps under ncu (one launch per case) for
# DRAM-throughput and tensor-pipe columns.  Usage: scripts/gpu_kernel_sweep.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
python scripts/kernel_sweep.py > $OUT/${TAG}_kernel_sweep.jsonl 2> $OUT/${TAG}_kernel_sweep.err; echo "kernel sweep rc=$?"
python scripts/conv_sweep.py > $OUT/${TAG}_conv_sweep.jsonl 2> $OUT/${TAG}_conv_sweep.err; echo "conv sweep rc=$?"
SIS_SWEEP_NCU=1 ncu --metrics $M --clock-control none -k regex:"fused_bias_act|upfirdn2d|label_" --csv --page raw --log-file $OUT/${TAG}_kernel_sweep_ncu.csv python scripts/kernel_sweep.py > $OUT/${TAG}_kernel_sweep_ncu_cases.jsonl 2>> $OUT/${TAG}_kernel_sweep.err; echo "ncu kernel sweep rc=$?"
SIS_SWEEP_NCU=1 ncu --metrics $M --clock-control none -k regex:"modconv_tc|blur_act_split" --csv --page raw --log-file $OUT/${TAG}_conv_sweep_ncu.csv python scripts/conv_sweep.py > $OUT/${TAG}_conv_sweep_ncu_cases.jsonl 2>> $OUT/${TAG}_conv_sweep.err; echo "ncu conv sweep rc=$?"
