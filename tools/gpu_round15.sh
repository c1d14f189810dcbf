#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== full pytest"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
for mode in pipeline graph; do
echo "== bench mode=$mode"
timeout 600 python bench.py --steps 30 --warmup 6 --mode $mode --cpu-frames 0 --extras 0 > gpurun_out/bench_$mode.log 2> gpurun_out/bench_$mode.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_$mode.log; tail -3 gpurun_out/bench_$mode.err
done
echo "== ncu launch list (graph)"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_tc.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --extras 0 --mode graph > gpurun_out/ncu_list_tc.log 2>&1; echo "rc=$?"
echo "== ncu full: 5x5 halo kernels (aggregator0, decoder block 4, refine)"
DVMVS_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_halo_kernel<32, 5" -c 8 -o gpurun_out/prof_halo_5x5 -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --extras 0 --mode eager > gpurun_out/ncu_full_halo.log 2>&1; echo "rc=$?"
du -sh gpurun_out
