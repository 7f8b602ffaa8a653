#!/bin/bash
# dataset leg on one GPU: native PNG writer (deflate, stored) and, for comparison, host contours.  Usage: scripts/gpu_leg_dataset.sh <tag>
TAG=${1:-r03}
for PNG in fast stored; do
  timeout 600 python bench.py --leg dataset --steps 150 --png $PNG > gpurun_out/${TAG}_leg_dataset_n1_$PNG.json 2> gpurun_out/${TAG}_leg_dataset_n1_$PNG.err; echo leg1_${PNG}_rc=$?
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_leg_dataset_n1_$PNG.json').read())
print({k:d[k] for k in ('value','pairs_generated','gpu_only_pairs_per_s','fraction_of_gpu_rate','png_encoder','host_cores','png_threads_per_rank','seconds_by_part','contour_stage','device_contour_stage')})"
  tail -3 gpurun_out/${TAG}_leg_dataset_n1_$PNG.err
done
timeout 600 python bench.py --leg dataset --steps 150 --contours host > gpurun_out/${TAG}_leg_dataset_n1_hostcontours.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_leg_dataset_n1_hostcontours.json').read()); print('host contours', d['value'], d['fraction_of_gpu_rate'])"
