#!/bin/bash
# session 2, call 11: plane sweep register cap / occupancy variants (roofline arm of bench.py)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for mb in 4 3 2; do
  DVMVS_SWEEP_MINB=$mb timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 20 2> gpurun_out/bench_minb$mb.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minb=$mb', round(d['value'],1), 'sweep_ms', round(d['roofline']['ms_per_launch'],4))"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "plane_sweep" 2>&1 | tail -2
