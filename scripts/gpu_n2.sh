#!/bin/bash
# 2-GPU check: bench under torchrun (both arms), the dataset leg (device contours; fast / stored PNG writer) at N = 1 and 2
TAG=${1:-r02}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo n2_rc=$?
tail -2 gpurun_out/${TAG}_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${TAG}_ref_n2.json 2> gpurun_out/${TAG}_ref_n2.err; echo ref_n2_rc=$?
for PNG in fast stored; do
  python bench.py --leg dataset --steps 40 --png $PNG > gpurun_out/${TAG}_leg_dataset_n1_$PNG.json 2> gpurun_out/${TAG}_leg_dataset_n1_$PNG.err; echo leg1_${PNG}_rc=$?; cat gpurun_out/${TAG}_leg_dataset_n1_$PNG.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --leg dataset --gpus 2 --steps 40 --png $PNG > gpurun_out/${TAG}_leg_dataset_n2_$PNG.json 2> gpurun_out/${TAG}_leg_dataset_n2_$PNG.err; echo leg2_${PNG}_rc=$?; cat gpurun_out/${TAG}_leg_dataset_n2_$PNG.json
done
python bench.py --leg contours > gpurun_out/${TAG}_leg_contours.json 2>/dev/null; cat gpurun_out/${TAG}_leg_contours.json
nproc
