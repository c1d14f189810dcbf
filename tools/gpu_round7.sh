#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench graph tc"
timeout 600 python bench.py --steps 20 --warmup 5 --mode graph --backend tc --cpu-frames 0 > gpurun_out/bench_graph_tc.log 2> gpurun_out/bench_graph_tc.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench_graph_tc.log; tail -2 gpurun_out/bench_graph_tc.err
echo "== tc_bench"
timeout 600 python tools/tc_bench.py 3 > gpurun_out/tc_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/tc_bench.log | cut -c1-330
echo "== ncu launch list (graph tc)"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_tc.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --mode graph --backend tc > gpurun_out/ncu_list_tc.log 2>&1; echo "rc=$?"
du -sh gpurun_out
