#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== full pytest"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 40 2> gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['config']['mode'], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches'])"
}
bench graph DVMVS_BENCH_MODE=graph
bench pipe5 A=1
bench pipe5_b A=1
echo "== stage times"
timeout 300 python tools/stage_times.py 5 2> gpurun_out/stage_times.err | tee gpurun_out/stage_times.json
du -sh gpurun_out
