"""Times the training-step kernels (row f3) alone at BASELINE config 5's per-GPU shapes (256x256, D=64, one measurement
frame per pair, subsequence of 8 => 7 pairs): plane-sweep forward / backward, gate epilogue forward / backward, multi-scale
loss forward / backward.  CUDA events on the current stream, L2 flushed between iterations.  Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import synth_data as synth  # noqa: E402
from dvmvs import training  # noqa: E402


def timed(fn, flush, n=20, warm=3):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def main():
    dev = torch.device("cuda", 0)
    B = int(os.environ.get("TRAIN_BENCH_BATCH", "4"))
    h = w = 128
    D = 64
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    f1 = cl(torch.randn(B, 32, h, w, device=dev)).requires_grad_(True)
    f2 = cl(torch.randn(B, 32, h, w, device=dev)).requires_grad_(True)
    pose1 = torch.from_numpy(np.stack([synth.camera_pose(1)] * B)).to(dev)
    pose2 = torch.from_numpy(np.stack([synth.camera_pose(0)] * B)).to(dev)
    K = synth.intrinsics(256, 256)
    K[0:2, :] /= 2.0
    K = torch.from_numpy(np.stack([K] * B)).to(dev)
    g = cl(torch.randn(B, D, h, w, device=dev))
    out = {"batch": B}

    def sweep_fwd():
        with torch.no_grad():
            training.plane_sweep_cost_volume(f1, [f2], pose1, [pose2], K, 0.25, 20.0, D)
    out["plane_sweep_forward_ms"] = timed(sweep_fwd, flush)
    cost = training.plane_sweep_cost_volume(f1, [f2], pose1, [pose2], K, 0.25, 20.0, D)

    def sweep_bwd():
        torch.autograd.grad(cost, [f1, f2], g, retain_graph=True)
    out["plane_sweep_backward_ms"] = timed(sweep_bwd, flush)
    # algorithmic bytes of the backward: read f1, f2, g; write g_f1; read-modify-write g_f2 (zero + accumulate counted once as a write)
    bytes_bwd = B * h * w * 4 * (32 + 32 + D + 32 + 32)
    out["plane_sweep_backward_algorithmic_GBps"] = bytes_bwd / (out["plane_sweep_backward_ms"] * 1e-3) / 1e9

    cc = cl(torch.randn(B, 2048, 8, 8, device=dev)).requires_grad_(True)
    c0 = cl(torch.randn(B, 512, 8, 8, device=dev)).requires_grad_(True)
    gh, gc = cl(torch.randn(B, 512, 8, 8, device=dev)), cl(torch.randn(B, 512, 8, 8, device=dev))

    def gates_fwd():
        with torch.no_grad():
            training.lstm_gate_epilogue(cc, c0)
    out["gate_epilogue_forward_ms"] = timed(gates_fwd, flush)
    hn, cn = training.lstm_gate_epilogue(cc, c0)

    def gates_bwd():
        torch.autograd.grad([hn, cn], [cc, c0], [gh, gc], retain_graph=True)
    out["gate_epilogue_backward_ms"] = timed(gates_bwd, flush)

    gt = torch.rand(B, 256, 256, device=dev) * 5 + 0.3
    preds = [(torch.rand(B, 256 // s, 256 // s, device=dev) * 5 + 0.3).requires_grad_(True) for s in (16, 8, 4, 2, 1)]

    def loss_fwd():
        with torch.no_grad():
            training.multi_scale_depth_loss(preds, [1, 1, 1, 1, 1], gt, "L1-inv")
    out["loss_forward_ms"] = timed(loss_fwd, flush)
    loss, _ = training.multi_scale_depth_loss(preds, [1, 1, 1, 1, 1], gt, "L1-inv")

    def loss_bwd():
        torch.autograd.grad(loss, preds, retain_graph=True)
    out["loss_backward_ms"] = timed(loss_bwd, flush)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
