#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
for cfg in "graph tc" "graph fp32" "eager tc"; do
  set -- $cfg
  echo "== bench mode=$1 backend=$2"
  timeout 600 python bench.py --steps 20 --warmup 5 --mode $1 --backend $2 --cpu-frames 0 > gpurun_out/bench_$1_$2.log 2> gpurun_out/bench_$1_$2.err; echo "bench exit $?"; cut -c1-1800 gpurun_out/bench_$1_$2.log; tail -3 gpurun_out/bench_$1_$2.err
done
echo "== ncu launch list (graph tc, profiler range)"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_tc.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --mode graph --backend tc > gpurun_out/ncu_list_tc.log 2>&1; echo "rc=$?"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_fp32.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --mode graph --backend fp32 > gpurun_out/ncu_list_fp32.log 2>&1; echo "rc=$?"
echo "== ncu full: plane sweep v2 + conv_tc"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plane_sweep_c32 -s 3 -c 1 -o gpurun_out/prof_sweep_v2 -f python bench.py --steps 2 --warmup 3 --cpu-frames 0 --mode eager --backend fp32 > gpurun_out/ncu_full_sweep.log 2>&1; echo "rc=$?"
DVMVS_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 60 -o gpurun_out/prof_conv_tc -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --mode eager --backend tc > gpurun_out/ncu_full_tc.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -40
