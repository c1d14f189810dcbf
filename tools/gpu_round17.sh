#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== full pytest (with shipped weights)"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "== ncu full: halo kernels of one keyframe"
DVMVS_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_halo_kernel -c 24 -o gpurun_out/prof_halo_final -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --extras 0 --mode eager > gpurun_out/ncu_full_halo.log 2>&1; echo "rc=$?"
du -sh gpurun_out
