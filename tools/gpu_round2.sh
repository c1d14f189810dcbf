#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== tc_check (first case alone, guards against hangs)"
timeout 120 python tools/tc_check.py k1_c64 > gpurun_out/tc_check_first.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/tc_check_first.log
if [ $rc -eq 0 ]; then
  timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; cat gpurun_out/tc_check.log | tail -20
  timeout 300 python tools/tc_e2e.py 3 > gpurun_out/tc_e2e.log 2>&1; echo "tc_e2e rc=$?"; tail -5 gpurun_out/tc_e2e.log
fi
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log | cut -c1-2500; tail -8 gpurun_out/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full: plane sweep"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plane_sweep_c32 -s 3 -c 2 -o gpurun_out/prof_sweep -f python bench.py --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
