#!/bin/bash
# session 2, call 8: tc default backend (smoke, full suite), tensor-pipe evidence summarised on the box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== smoke (library default backend)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== full pytest"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== ncu full: halo convs (fp16 operands), one keyframe"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:conv_halo_kernel -c 14 -o /tmp/prof_halo_fp16 -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --extras 0 --mode eager > gpurun_out/ncu_full_halo.log 2>&1; echo "rc=$?"
python tools/summarize_ncu.py table /tmp/prof_halo_fp16.ncu-rep gpurun_out/halo_fp16_table.md
python tools/summarize_ncu.py full /tmp/prof_halo_fp16.ncu-rep gpurun_out/halo_fp16_full.md
echo "== ncu full: conv_tc (fp16 operands), first 24 of a keyframe"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc_kernel -c 24 -o /tmp/prof_tc_fp16 -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --extras 0 --mode eager > gpurun_out/ncu_full_tc.log 2>&1; echo "rc=$?"
python tools/summarize_ncu.py table /tmp/prof_tc_fp16.ncu-rep gpurun_out/tc_fp16_table.md
cat gpurun_out/halo_fp16_table.md
du -sh gpurun_out
