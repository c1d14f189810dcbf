"""Event timings of the plane-sweep kernels alone (roofline arm of bench.py, more variants and batch sizes).

    python tools/sweep_bench.py [--clips 1,8,32] [--variants fp32,h16,tc1,tc3] [--out gpurun_out/sweep_bench.json]

c2 geometry (128x128 half-resolution features, 64 planes, 2 measurement frames, bench.py's synthetic poses); every
variant is checked against the fp32 fused kernel (max abs error relative to the cost volume's max magnitude) and timed
with CUDA events on the launching stream, L2 flushed between launches."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "deep-video-mvs_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", default="1,8,32")
    ap.add_argument("--variants", default="fp32,h16,tc1,tc3")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--hw", default="128,128")
    ap.add_argument("--planes", type=int, default=64)
    ap.add_argument("--meas", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import synth_data as synth
    from dvmvs import _ops as ops
    dev = torch.device("cuda", 0)
    h, w = [int(v) for v in args.hw.split(",")]
    D, M = args.planes, args.meas
    peaks = {"hbm_gbs": 6650.0}
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peaks = json.load(open(pk))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    results = []
    for B in [int(v) for v in args.clips.split(",")]:
        g = torch.Generator(device="cpu").manual_seed(B)
        f1 = (torch.randn(B, h, w, 32, generator=g) * 4).to(dev)
        f2 = [(torch.randn(B, h, w, 32, generator=g) * 4).to(dev) for _ in range(M)]
        K = torch.from_numpy(synth.intrinsics(2 * h, 2 * w))[None].repeat(B, 1, 1).to(dev)
        K[:, 0:2, :] /= 2.0
        pose1 = torch.from_numpy(synth.camera_pose(M))[None].repeat(B, 1, 1).to(dev)
        pose2 = [torch.from_numpy(synth.camera_pose(M - k))[None].repeat(B, 1, 1).to(dev) for k in range(1, M + 1)]
        bytes_alg = ((1 + M) * 32 + D) * h * w * 4 * B
        base = ops.plane_sweep(f1, f2, pose1, pose2, K, 0.25, 20.0, D, True)
        scale = float(base.abs().max())
        p1 = ops.split_planes(f1)
        p2 = [ops.split_planes(t) for t in f2]

        def run(variant):
            if variant == "fp32":
                return ops.plane_sweep(f1, f2, pose1, pose2, K, 0.25, 20.0, D, True)
            if variant == "h16":
                return ops.plane_sweep_h16(f1, [p[0] for p in p2], pose1, pose2, K, 0.25, 20.0, D)
            if variant in ("tc1", "tc3"):
                return ops.plane_sweep_tc(p1, p2, pose1, pose2, K, 0.25, 20.0, D, terms=int(variant[2]))
            raise ValueError(variant)

        for variant in args.variants.split(","):
            if variant.startswith("tc") and not hasattr(ops, "plane_sweep_tc"):
                continue
            try:
                out = run(variant)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                results.append({"clips": B, "variant": variant, "error": str(e)[:300]})
                print(results[-1], flush=True)
                continue
            err = float((out - base).abs().max()) / scale
            for _ in range(3):
                run(variant)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
            for a, b in evs:
                flush.zero_()
                a.record()
                run(variant)
                b.record()
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
            gbs = bytes_alg / (ms * 1e-3) / 1e9
            results.append({"clips": B, "variant": variant, "ms": ms, "us_per_clip": ms * 1e3 / B, "algorithmic_GBps": gbs,
                            "frac_of_measured_hbm": gbs / peaks["hbm_gbs"], "rel_err_vs_fp32_kernel": err,
                            "h": h, "w": w, "D": D, "M": M})
            print(results[-1], flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
