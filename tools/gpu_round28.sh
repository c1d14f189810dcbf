#!/bin/bash
# session 2, call 5: host enqueue cost, launch list with fp16 operands, engine variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 40 2> gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['config']['mode'], round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'host_enq_ms', d['config'].get('host_enqueue_ms_per_step'), 'sweep', round(d['roofline']['ms_per_launch'],4))"
}
bench pipe5 A=1
bench pipe5_prio DVMVS_PIPE_PRIO=1
bench pipe5_prio_nopdl DVMVS_PIPE_PRIO=1 DVMVS_PIPE_REC_PDL=0
timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 40 --stages 3 2> gpurun_out/bench_s3.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stages3', round(d['value'],1), d['config'].get('host_enqueue_ms_per_step'))"
timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 40 --stages 4 2> gpurun_out/bench_s4.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stages4', round(d['value'],1), d['config'].get('host_enqueue_ms_per_step'))"
echo "== stage times (fp16 operands)"
DVMVS_TC_TERMS=1 timeout 300 python tools/stage_times.py 5 2> gpurun_out/stage_times.err | tee gpurun_out/stage_times.json
echo "== ncu launch list (graph, fp16 operands)"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_fp16.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --extras 0 --mode graph > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
