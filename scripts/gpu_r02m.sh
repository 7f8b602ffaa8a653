#!/bin/bash
# contour stage on the device: pipeline tests + dataset leg with host / device contours (1 GPU)
TAG=${1:-r02m}
if [ "$2" != "notest" ]; then timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_contours_gpu.py -x -q -m gpu 2>&1 | tail -8; fi
for MODE in host device; do
  timeout 600 python bench.py --leg dataset --steps 40 --contours $MODE > gpurun_out/${TAG}_leg_dataset_${MODE}.json 2> gpurun_out/${TAG}_leg_dataset_${MODE}.err; echo leg_${MODE}_rc=$?
  cat gpurun_out/${TAG}_leg_dataset_${MODE}.json; tail -3 gpurun_out/${TAG}_leg_dataset_${MODE}.err
done
nproc
