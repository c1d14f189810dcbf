#!/bin/bash
# session 2, call 7: online engine test, tensor-pipe evidence for the fp16-operand conv kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== online engine + feature cache tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -k "online or feature_cache or shipped" 2>&1 | tail -6
echo "== ncu full: halo convs (fp16 operands), one keyframe"
DVMVS_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_halo_kernel -c 14 -o gpurun_out/prof_halo_fp16 -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --extras 0 --mode eager > gpurun_out/ncu_full_halo.log 2>&1; echo "rc=$?"
echo "== ncu full: conv_tc (fp16 operands), one keyframe"
DVMVS_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc_kernel -c 60 -o gpurun_out/prof_tc_fp16 -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --extras 0 --mode eager > gpurun_out/ncu_full_tc.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
