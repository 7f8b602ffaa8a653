#!/bin/bash
# N-GPU check of the bench and the dataset leg.  Usage: scripts/gpu_n4.sh <tag> <N>
TAG=${1:-r03}; N=${2:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --extra-configs= --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo bench_rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --leg dataset --gpus $N --steps 150 > gpurun_out/${TAG}_leg_dataset_n$N.json 2> gpurun_out/${TAG}_leg_dataset_n$N.err; echo leg_rc=$?; cut -c1-700 gpurun_out/${TAG}_leg_dataset_n$N.json
nproc
