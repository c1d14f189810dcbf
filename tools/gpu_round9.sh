#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
nvidia-smi -L
echo "== pytest -m gpu (subset: modules + stem)"
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x -k "modules or baseline_configs or pipeline" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench 1 GPU"
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-frames 0 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_n1.log
echo "== bench 2 GPUs (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-frames 0 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err
echo "== reference arm under torchrun"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_ref_n2.log 2> gpurun_out/bench_ref_n2.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref_n2.log
