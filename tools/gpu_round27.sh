#!/bin/bash
# session 2, call 4: fp16-operand default + precision-parametrised tests, sweep L1 prefetch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "== full pytest"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== sweep: prefetch off / on (bench roofline arm)"
for pf in 0 1; do
  DVMVS_SWEEP_PREFETCH=$pf timeout 300 python bench.py --cpu-frames 0 --extras 0 --steps 20 2> gpurun_out/bench_pf$pf.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prefetch=$pf', round(d['value'],1), 'sweep_ms', round(d['roofline']['ms_per_launch'],4))"
done
echo "== bench default"
timeout 600 python bench.py 2> gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['ms_per_launch'], d['operating_points'])"
tail -3 gpurun_out/bench_default.err
echo "== ncu full: sweep"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plane_sweep_c32 -s 3 -c 1 -o gpurun_out/prof_sweep_v6 -f python bench.py --steps 2 --warmup 3 --cpu-frames 0 --extras 0 --mode eager > gpurun_out/ncu_full_sweep.log 2>&1; echo "rc=$?"
