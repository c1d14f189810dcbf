"""Steady-state timing (CUDA events, warm caches, back-to-back launches) of individual convolution layers on both
backends, for the layer shapes that dominate the keyframe.  Usage: python tools/tc_bench.py [terms]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import torch

import synth_data as synth
from dvmvs import _native as N
from dvmvs import _ops as ops

DEV = "cuda"
LAYERS = [
    # name, B, H, W, [src channels], Cout, k, stride
    ("refine.1 5x5 32->32 @256^2", 1, 256, 256, [32], 32, 5, 1),
    ("refine.0 5x5 36->32 @256^2 (packed)", 1, 256, 256, [36], 32, 5, 1),
    ("db4.conv1 5x5 65->32 @128^2", 1, 128, 128, [32, 32, 1], 32, 5, 1),
    ("aggregator0 5x5 96->32 @128^2", 1, 128, 128, [32, 64], 32, 5, 1),
    ("fpn.layer0 3x3 32->32 @128^2 B=3", 3, 128, 128, [32], 32, 3, 1),
    ("eb0.conv 5x5 64->64 @64^2", 1, 64, 64, [64], 64, 5, 1),
    ("eb1.conv 3x3 128->128 @32^2", 1, 32, 32, [128], 128, 3, 1),
    ("eb2.conv 3x3 256->256 @16^2", 1, 16, 16, [256], 256, 3, 1),
    ("eb3.conv 3x3 512->512 @8^2", 1, 8, 8, [512], 512, 3, 1),
    ("lstm 3x3 1024->2048 @8^2", 1, 8, 8, [512, 512], 2048, 3, 1),
    ("mnas pw 1x1 96->576 @16^2 B=3", 3, 16, 16, [96], 576, 1, 1),
    ("mnas pw 1x1 1152->192 @8^2 B=3", 3, 8, 8, [1152], 192, 1, 1),
    ("mnas pw 1x1 16->48 @128^2 B=3", 3, 128, 128, [16], 48, 1, 1),
]


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    terms = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    for name, B, H, W, chans, Cout, k, stride in LAYERS:
        cin = sum(chans)
        xs = [torch.from_numpy(synth.tensor("tb/x%d" % i, (B, H, W, c), seed=1)).to(DEV) for i, c in enumerate(chans)]
        w = torch.from_numpy(synth.tensor("tb/w", (Cout, cin, k, k), seed=2, scale=(2.0 / (cin * k * k)) ** 0.5))
        pc = ops.PackedConv(w, None, None, stride=stride, act=N.ACT_RELU)
        ptc = ops.PackedConvTC(pc, chans, DEV)
        pc.weight = pc.weight.to(DEV)
        planes = [ops.split_planes(x) for x in xs]
        t_fp32 = timeit(lambda: ops.conv2d([(x, N.SRC_DIRECT) for x in xs], pc))
        res = []
        for bn in (32, 64, 128):
            if bn > 32 and Cout <= 32:
                continue
            for split in (False, True):
                t = timeit(lambda: ops.conv2d_tc(planes, ptc, terms=terms, block_n=bn, allow_split=split))
                res.append("N%d%s %.1f" % (bn, "+splitK" if split else "", t))
        t1 = timeit(lambda: ops.conv2d_tc(planes, ptc, terms=1, allow_split=True))
        if stride == 1 and k >= 3 and Cout <= 64 * 8:
            for kc in (16, 32):
                try:
                    ph = ops.PackedConvHalo(pc, chans, DEV, kc=kc, concat_padded=True)
                    blk = ops.split_blocked([(x, False) for x in xs])
                    th = timeit(lambda: ops.conv2d_halo([blk], ph, terms=terms, want_f32=True, want_blk=True, want_nhwc=False))
                    res.append("HALO kc%d %.1f" % (kc, th))
                except Exception as e:  # noqa: BLE001
                    res.append("HALO kc%d FAILED %s" % (kc, str(e)[:60]))
        macs = B * (H // stride) * (W // stride) * Cout * cin * k * k
        print("%-38s %7.1f MMAC | fp32 %7.1f us | tc x%d: %s | tc x1 auto %.1f us" % (name, macs / 1e6, t_fp32, terms, "  ".join(res), t1), flush=True)


if __name__ == "__main__":
    main()
