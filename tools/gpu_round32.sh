#!/bin/bash
# session 2, call 9: fused split-K finish (last-arriving CTA reduces) -- parity + A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== tensor-core tests (fused finish default, separate finish in the child-process test)"
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -k "conv2d_tc or tensor_core or split_k or pipeline_keyframe or finish" 2>&1 | tail -6
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 60 2> gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['config']['mode'], round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])"
}
bench fused A=1
bench separate DVMVS_SPLITK_FUSED=0
bench fused_b A=1
bench separate_b DVMVS_SPLITK_FUSED=0
bench fused_graph DVMVS_BENCH_MODE=graph
bench separate_graph DVMVS_BENCH_MODE=graph DVMVS_SPLITK_FUSED=0
