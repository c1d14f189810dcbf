#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== tc + halo unit tests"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider --timeout 60 -k "halo or conv2d_tc" > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_tc.log | cut -c1-300
echo "== timeline"
timeout 120 python tools/halo_timeline.py 2>&1 | cut -c1-200
echo "== tc_bench"
timeout 600 python tools/tc_bench.py 3 > gpurun_out/tc_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/tc_bench.log | cut -c1-420
echo "== full pytest"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
for mode in pipeline graph; do
echo "== bench mode=$mode"
timeout 600 python bench.py --steps 30 --warmup 6 --mode $mode --cpu-frames 0 --extras 0 > gpurun_out/bench_$mode.log 2> gpurun_out/bench_$mode.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_$mode.log; tail -3 gpurun_out/bench_$mode.err
done
echo "== bench pipeline, halo off"
DVMVS_HALO=0 timeout 600 python bench.py --steps 30 --warmup 6 --mode pipeline --cpu-frames 0 --extras 0 > gpurun_out/bench_pipeline_nohalo.log 2>&1; cut -c1-330 gpurun_out/bench_pipeline_nohalo.log | tail -1
