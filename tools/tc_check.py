"""GPU check of the tcgen05 implicit-GEMM convolution against the fp32 CUDA-core kernel and torch fp32 (CPU).
Prints one line per case; run on the GPU box (tools/gpu_round.sh)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import numpy as np
import torch
import torch.nn.functional as F

import synth_data as synth
from dvmvs import _native as N
from dvmvs import _ops as ops

DEV = "cuda"
CASES = [
    # name, B, H, W, [src real channels], Cout, k, stride, act, block_n, terms
    ("k1_c64", 1, 16, 16, [64], 32, 1, 1, 0, 32, 3),
    ("k1_c32", 1, 16, 16, [32], 32, 1, 1, 0, 32, 3),
    ("k3_c64", 1, 16, 16, [64], 32, 3, 1, 1, 32, 3),
    ("k3_c32_n64", 1, 32, 32, [32], 64, 3, 1, 1, 64, 3),
    ("k5_agg0", 1, 64, 64, [32, 64], 32, 5, 1, 1, 32, 3),
    ("k3_odd", 2, 24, 40, [32, 128], 128, 3, 1, 1, 128, 3),
    ("k3_small", 1, 8, 10, [512, 512], 96, 3, 1, 0, 32, 3),
    ("k3_dec", 1, 16, 16, [128, 128, 1], 128, 3, 1, 1, 64, 3),
    ("k5_refine", 1, 32, 32, [32, 1, 3], 32, 5, 1, 1, 32, 3),
    ("k3_fp16x1", 1, 32, 32, [64], 64, 3, 1, 1, 64, 1),
    ("k3_s2", 1, 32, 32, [64], 128, 3, 2, 1, 64, 3),
    ("k5_s2", 1, 64, 64, [32], 64, 5, 2, 1, 64, 3),
]


def main():
    only = sys.argv[1:] or None
    for (name, B, H, W, chans, Cout, k, stride, act, block_n, terms) in CASES:
        if only and name not in only:
            continue
        cin = sum(chans)
        xs = [torch.from_numpy(synth.tensor("tc/%s/x%d" % (name, i), (B, H, W, c), seed=1)).to(DEV) for i, c in enumerate(chans)]
        w = torch.from_numpy(synth.tensor("tc/%s/w" % name, (Cout, cin, k, k), seed=2, scale=(2.0 / (cin * k * k)) ** 0.5))
        bias = torch.from_numpy(synth.tensor("tc/%s/b" % name, (Cout,), seed=3, scale=0.1))
        pc = ops.PackedConv(w, bias, None, stride=stride, act=act)
        ptc = ops.PackedConvTC(pc, chans, DEV)
        pc.weight, pc.bias = pc.weight.to(DEV), pc.bias.to(DEV)
        ref = ops.conv2d([(x, N.SRC_DIRECT) for x in xs], pc)
        planes = [ops.split_planes(x) for x in xs]
        try:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out_f32, out_planes = ops.conv2d_tc(planes, ptc, terms=terms, block_n=block_n, allow_split=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            rec = out_planes[0].float() + out_planes[1].float()
            scale = float(ref.abs().max())
            e32 = float((out_f32 - ref).abs().max()) / scale
            epl = float((rec - ref).abs().max()) / scale
            # split-K variant
            o2, _ = ops.conv2d_tc(planes, ptc, terms=terms, block_n=block_n, allow_split=True)
            torch.cuda.synchronize()
            es = float((o2 - ref).abs().max()) / scale
            print("%-12s max|tc-fp32|/max = %.3e  planes %.3e  splitK %.3e  (%.2f ms first call)" % (name, e32, epl, es, dt * 1e3), flush=True)
        except Exception as e:  # noqa: BLE001
            print("%-12s FAILED: %s" % (name, e), flush=True)
            break


if __name__ == "__main__":
    main()
