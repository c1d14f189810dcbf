#!/bin/bash
# session 2, call 2: training-step kernels (row f3)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== training tests"
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu_training.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu_training.log | cut -c1-400
echo "== training kernel timings"
timeout 300 python tools/train_bench.py 2> gpurun_out/train_bench.err | tee gpurun_out/train_bench.json
tail -5 gpurun_out/train_bench.err
