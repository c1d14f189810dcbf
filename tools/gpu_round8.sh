#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== pytest -m gpu (PDL on)"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
for pdl in 1 0; do
echo "== bench graph tc PDL=$pdl"
DVMVS_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --mode graph --backend tc --cpu-frames 0 > gpurun_out/bench_graph_tc_pdl$pdl.log 2> gpurun_out/bench_graph_tc_pdl$pdl.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench_graph_tc_pdl$pdl.log; tail -2 gpurun_out/bench_graph_tc_pdl$pdl.err
done
echo "== bench eager tc PDL=1"
DVMVS_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --mode eager --backend tc --cpu-frames 0 > gpurun_out/bench_eager_tc_pdl1.log 2>&1; echo "bench exit $?"; cut -c1-330 gpurun_out/bench_eager_tc_pdl1.log
du -sh gpurun_out
