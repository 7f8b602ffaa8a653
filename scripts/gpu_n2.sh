#!/bin/bash
# 2-GPU check: bench under torchrun (both arms; configs 2 and 4), the dataset leg (device contours, native PNG writer) at N = 1 and 2
TAG=${1:-r02}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo n2_rc=$?
tail -2 gpurun_out/${TAG}_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${TAG}_ref_n2.json 2> gpurun_out/${TAG}_ref_n2.err; echo ref_n2_rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --config 4 --extra-configs= --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_cfg4_n2.json 2> gpurun_out/${TAG}_bench_cfg4_n2.err; echo cfg4_n2_rc=$?
python bench.py --leg dataset --steps 150 > gpurun_out/${TAG}_leg_dataset_n1.json 2> gpurun_out/${TAG}_leg_dataset_n1.err; echo leg1_rc=$?; cut -c1-600 gpurun_out/${TAG}_leg_dataset_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --leg dataset --gpus 2 --steps 150 > gpurun_out/${TAG}_leg_dataset_n2.json 2> gpurun_out/${TAG}_leg_dataset_n2.err; echo leg2_rc=$?; cut -c1-600 gpurun_out/${TAG}_leg_dataset_n2.json
nproc
