#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== pipeline test"
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -k "pipeline" > gpurun_out/pytest_pipe.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_pipe.log | cut -c1-300
echo "== stage times"
timeout 300 python tools/stage_times.py 3 4 5 2> gpurun_out/stage_times.err | tee gpurun_out/stage_times.json
for st in 3 4 5; do for pr in 0 1; do
echo "== bench stages=$st prio=$pr"
DVMVS_PIPE_PRIO=$pr timeout 300 python bench.py --stages $st --cpu-frames 0 --extras 0 --steps 40 2> gpurun_out/bench_s${st}_p${pr}.err | tee gpurun_out/bench_s${st}_p${pr}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done
du -sh gpurun_out
