#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== halo unit tests (first alone: hang guard)"
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 60 -x -k "halo and k3_c32" > gpurun_out/pytest_halo_first.log 2>&1; rc=$?
echo "rc=$rc"; tail -12 gpurun_out/pytest_halo_first.log | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider --timeout 60 -k "halo" > gpurun_out/pytest_halo.log 2>&1; echo "rc=$?"; tail -14 gpurun_out/pytest_halo.log | cut -c1-400
echo "== full pytest"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== tc_bench"
timeout 600 python tools/tc_bench.py 3 > gpurun_out/tc_bench.log 2>&1; echo "rc=$?"; head -8 gpurun_out/tc_bench.log | cut -c1-400
for mode in pipeline graph; do
echo "== bench mode=$mode"
timeout 600 python bench.py --steps 30 --warmup 6 --mode $mode --cpu-frames 0 > gpurun_out/bench_$mode.log 2> gpurun_out/bench_$mode.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_$mode.log; tail -3 gpurun_out/bench_$mode.err
done
