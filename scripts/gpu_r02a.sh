#!/bin/bash
# round-2 check A: GPU tests, default bench (config 2 + extra configs 3, 4), reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo pytest_rc=$?; tail -5 gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo bench_rc=$?; tail -3 gpurun_out/r02a_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err; echo ref_rc=$?
nproc; free -g | head -2
