#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
echo "== bench 1 GPU (default flags)"
timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; cat gpurun_out/bench_n1.log | cut -c1-3000; tail -4 gpurun_out/bench_n1.err
echo "== bench 2 GPUs (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-frames 0 --extras 0 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err
