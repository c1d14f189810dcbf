#!/bin/bash
# session 2, call 3: operand-precision policies (accuracy + throughput), sweep with predicated out-of-image taps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== sweep parity (predicated taps)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -k "plane_sweep or baseline_configs" 2>&1 | tail -4
echo "== terms probe"
timeout 600 python tools/terms_probe.py 2> gpurun_out/terms_probe.err | tee gpurun_out/terms_probe.jsonl | cut -c1-600
tail -3 gpurun_out/terms_probe.err
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 40 2> gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches'], 'sweep_ms', round(d['roofline']['ms_per_launch'],4))"
}
bench all3 A=1
bench fe_fpn_cve DVMVS_TC_POLICY=fe=1,fpn=1,cve=1
bench no_cvd DVMVS_TC_POLICY=fe=1,fpn=1,cve=1,lstm=1
bench all1 DVMVS_TC_POLICY=fe=1,fpn=1,cve=1,lstm=1,cvd=1
bench all1_graph DVMVS_TC_POLICY=fe=1,fpn=1,cve=1,lstm=1,cvd=1 DVMVS_BENCH_MODE=graph
