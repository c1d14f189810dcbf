#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
for mode in pipeline graph; do
echo "== bench mode=$mode"
timeout 600 python bench.py --steps 30 --warmup 6 --mode $mode --cpu-frames 0 > gpurun_out/bench_$mode.log 2> gpurun_out/bench_$mode.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_$mode.log; tail -3 gpurun_out/bench_$mode.err
done
echo "== bench mode=pipeline clips=4"
timeout 600 python bench.py --steps 30 --warmup 6 --mode pipeline --clips 4 --cpu-frames 0 > gpurun_out/bench_pipeline_b4.log 2> gpurun_out/bench_pipeline_b4.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_pipeline_b4.log; tail -3 gpurun_out/bench_pipeline_b4.err
