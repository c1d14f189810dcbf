#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== full pytest (with shipped weights)"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
for st in 3 2; do
echo "== bench pipeline stages=$st"
timeout 600 python bench.py --steps 40 --warmup 8 --mode pipeline --stages $st --cpu-frames 0 --extras 0 > gpurun_out/bench_pipeline_s$st.log 2> gpurun_out/bench_pipeline_s$st.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_pipeline_s$st.log; tail -3 gpurun_out/bench_pipeline_s$st.err
done
echo "== bench default (all arms)"
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "rc=$?"; cut -c1-2600 gpurun_out/bench_default.log; tail -3 gpurun_out/bench_default.err
