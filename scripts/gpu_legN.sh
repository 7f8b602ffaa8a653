#!/bin/bash
# dataset leg on N GPUs.  Usage: scripts/gpu_legN.sh <tag> <N>
TAG=${1:-r03}; N=${2:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --leg dataset --gpus $N --steps 150 > gpurun_out/${TAG}_leg_dataset_n$N.json 2> gpurun_out/${TAG}_leg_dataset_n$N.err; echo leg_rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_leg_dataset_n$N.json').read())
print({k:d[k] for k in ('value','gpu_only_pairs_per_s','fraction_of_gpu_rate','host_cores','png_threads_per_rank','contour_workers_per_rank','seconds_by_part','contour_stage')})"
tail -2 gpurun_out/${TAG}_leg_dataset_n$N.err
