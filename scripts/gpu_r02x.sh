#!/bin/bash
TAG=${1:-r02x}
timeout 600 python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -4
for PNG in fast stored cv2; do
  timeout 600 python bench.py --leg dataset --steps 40 --png $PNG > gpurun_out/${TAG}_leg_dataset_n1_$PNG.json 2> gpurun_out/${TAG}_leg_dataset_n1_$PNG.err; echo leg1_${PNG}_rc=$?
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_leg_dataset_n1_$PNG.json').read())
print({k:d[k] for k in ('value','gpu_only_pairs_per_s','fraction_of_gpu_rate','png_encoder','host_cores','png_threads_per_rank','seconds_by_part','contour_stage')})"
done
