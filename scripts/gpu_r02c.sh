#!/bin/bash
# round-2 check C: GPU tests + default bench after the whole-warp issue loops / UP4 / BK=32 halo kernels
TAG=${1:-r02c}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo pytest_rc=$?; tail -5 gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench_rc=$?; tail -3 gpurun_out/${TAG}_bench.err
python - <<'P'
import json,sys
d=json.loads(open('gpurun_out/'+sys.argv[1]+'_bench.json').read()) if len(sys.argv)>1 else None
P
