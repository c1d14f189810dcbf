#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench graph tc"
timeout 600 python bench.py --steps 20 --warmup 5 --mode graph --backend tc --cpu-frames 0 > gpurun_out/bench_graph_tc.log 2> gpurun_out/bench_graph_tc.err; echo "bench exit $?"; cut -c1-700 gpurun_out/bench_graph_tc.log; tail -3 gpurun_out/bench_graph_tc.err
echo "== bench graph tc, 8 clips"
timeout 600 python bench.py --steps 20 --warmup 5 --mode graph --backend tc --cpu-frames 0 --clips 8 > gpurun_out/bench_graph_tc_b8.log 2> gpurun_out/bench_graph_tc_b8.err; echo "bench exit $?"; cut -c1-700 gpurun_out/bench_graph_tc_b8.log; tail -3 gpurun_out/bench_graph_tc_b8.err
echo "== ncu launch list (graph tc)"
DVMVS_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_tc.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --mode graph --backend tc > gpurun_out/ncu_list_tc.log 2>&1; echo "rc=$?"
echo "== ncu full: last 2 conv_tc launches of a keyframe (refine.0, refine.1) + aggregator0"
N=$(python - <<'PY'
import csv, io, re
t = open("gpurun_out/launches_graph_tc.csv").read()
rows = list(csv.DictReader(io.StringIO(t[t.find('"ID"'):])))
n = sum(1 for r in rows if "conv_tc_kernel" in r["Kernel Name"]) // 2
print(n)
PY
)
echo "conv_tc launches per keyframe: $N"
DVMVS_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $((N-2)) -c 2 -o gpurun_out/prof_conv_tc_refine -f python bench.py --steps 1 --warmup 3 --cpu-frames 0 --mode eager --backend tc > gpurun_out/ncu_full_tc.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -40; du -sh gpurun_out
