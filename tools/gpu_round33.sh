#!/bin/bash
# session 2, call 10: validation of the final state -- full GPU suite, smoke, default bench record
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== full pytest"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench default"
timeout 900 python bench.py 2> gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['ms_per_launch'], d['clocks'], d['operating_points'])"
tail -2 gpurun_out/bench_default.err
