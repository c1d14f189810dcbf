#!/bin/bash
# session 2, call 6: default bench record (100 steps, c3 operating point), smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench default"
timeout 900 python bench.py 2> gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['ms_per_launch'], d['clocks'], d['operating_points'])"
tail -3 gpurun_out/bench_default.err
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_reference.err | tee gpurun_out/bench_reference.json | cut -c1-600
