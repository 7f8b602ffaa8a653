#!/bin/bash
TAG=${1:-r02x}
timeout 600 python -m pytest tests/test_contours_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python scripts/contour_stage_bench.py > gpurun_out/${TAG}_contour_stage.jsonl 2> gpurun_out/${TAG}_contour_stage.err; cat gpurun_out/${TAG}_contour_stage.jsonl; tail -3 gpurun_out/${TAG}_contour_stage.err
for PNG in fast stored; do
  timeout 600 python bench.py --leg dataset --steps 40 --png $PNG > gpurun_out/${TAG}_leg_dataset_n1_$PNG.json 2> gpurun_out/${TAG}_leg_dataset_n1_$PNG.err; echo leg1_${PNG}_rc=$?
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_leg_dataset_n1_$PNG.json').read())
print({k:d[k] for k in ('value','gpu_only_pairs_per_s','fraction_of_gpu_rate','png_encoder','host_cores','png_threads_per_rank','seconds_by_part','contour_stage','device_contour_stage')})"
  tail -3 gpurun_out/${TAG}_leg_dataset_n1_$PNG.err
done
