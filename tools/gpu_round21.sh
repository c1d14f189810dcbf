#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 40 2> gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['config']['mode'], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches'])"
}
bench base A=1
bench prio_nopdl DVMVS_PIPE_PRIO=1 DVMVS_PIPE_REC_PDL=0
bench nopdl_rec DVMVS_PIPE_REC_PDL=0
bench prio DVMVS_PIPE_PRIO=1
bench sweep3 DVMVS_SWEEP_CTAS_PER_SM=3
bench sweep2 DVMVS_SWEEP_CTAS_PER_SM=2
bench base2 A=1
echo "== launch list (graph mode)"
DVMVS_PROFILE=1 DVMVS_BENCH_MODE=graph timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_graph_r21.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --extras 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== default bench (full, with extras + cpu baseline)"
timeout 900 python bench.py 2> gpurun_out/bench_default.err | tee gpurun_out/bench_default.log | cut -c1-600
du -sh gpurun_out
