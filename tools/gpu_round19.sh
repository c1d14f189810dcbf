#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
run_tests() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 120 -x -k "conv2d_tc or conv2d_halo or tensor_core_backend_vs_oracle or conv2d_vs_torch or preprocessing" > gpurun_out/pytest_$name.log 2>&1
  echo "== tests[$name] exit $?"; tail -4 gpurun_out/pytest_$name.log | cut -c1-400
}
run_tests all_on A=1
run_tests cluster_on DVMVS_CLUSTER_SPLITK=1
run_tests cat_off DVMVS_TC_CAT=0 DVMVS_HALO_CAT=0
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 40 2> gpurun_out/bench_$name.err | tee gpurun_out/bench_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['config']['mode'], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches'])"
}
bench graph_all_on DVMVS_BENCH_MODE=graph
bench graph_cluster_off DVMVS_BENCH_MODE=graph DVMVS_CLUSTER_SPLITK=0
bench graph_cat_off DVMVS_BENCH_MODE=graph DVMVS_TC_CAT=0 DVMVS_HALO_CAT=0
bench graph_halocat_off DVMVS_BENCH_MODE=graph DVMVS_HALO_CAT=0
bench pipe5_all_on A=1
bench pipe5_all_off DVMVS_CLUSTER_SPLITK=0 DVMVS_TC_CAT=0 DVMVS_HALO_CAT=0
echo "== full pytest"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
du -sh gpurun_out
