#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
du -sh tests/golden/_ref_data 2>/dev/null
echo "== ncu launch lists (profiler range = 2 timed keyframes)"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_tc.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --mode graph --backend tc > gpurun_out/ncu_list_tc.log 2>&1; echo "rc=$?"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_fp32.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --mode graph --backend fp32 > gpurun_out/ncu_list_fp32.log 2>&1; echo "rc=$?"
echo "== ncu full: plane sweep v2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plane_sweep_c32 -s 3 -c 1 -o gpurun_out/prof_sweep_v2 -f python bench.py --steps 2 --warmup 3 --cpu-frames 0 --mode eager --backend fp32 > gpurun_out/ncu_full_sweep.log 2>&1; echo "rc=$?"
echo "== ncu full: 6 conv_tc launches (refine-sized ones are the last of a keyframe)"
DVMVS_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc_kernel -s 52 -c 6 -o gpurun_out/prof_conv_tc -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --mode eager --backend tc > gpurun_out/ncu_full_tc.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -40; du -sh gpurun_out
