"""Golden-vector case definitions shared by oracle/make_golden.py (reference side, build container) and the
tests (oracle side / CUDA side).  numpy only; `synth` is deep-video-mvs_b200/synth_data.py passed in by the
caller (this file is loaded by path from two different processes).  Inputs are regenerated from seeds on both
sides; only reference OUTPUTS are stored in ops.npz / modules.npz.
"""
import numpy as np


def _pose(rng, trans_scale, rot_scale):
    """Random rigid cam-to-world pose: small rotation (Rodrigues) + translation."""
    ax = rng.randn(3)
    ax /= np.linalg.norm(ax)
    ang = rot_scale * rng.uniform(0.3, 1.0)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    P = np.eye(4)
    P[:3, :3] = R
    P[:3, 3] = trans_scale * rng.uniform(-1, 1, size=3)
    return P.astype(np.float32)


def _K(fx, fy, cx, cy):
    return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32)


# ---------------------------------------------------------------------------------------------- plane sweep
# dot: C=32 features (the fusionnet/pairnet use); sad: C=3 raw-RGB branch (utils.py:83-84, baselines only).
PLANE_SWEEP_CASES = {
    # name: B, C, h, w, D, M, dot, pose scales
    "dot_small":   dict(B=2, C=32, h=24, w=40, D=16, M=2, dot=True,  trans=0.15, rot=0.05, min_depth=0.25, max_depth=20.0, seed=1),
    "dot_c1":      dict(B=1, C=32, h=64, w=64, D=32, M=1, dot=True,  trans=0.12, rot=0.03, min_depth=0.25, max_depth=20.0, seed=2),
    "dot_m3":      dict(B=1, C=32, h=32, w=40, D=64, M=3, dot=True,  trans=0.20, rot=0.10, min_depth=0.25, max_depth=20.0, seed=3),
    "dot_wide":    dict(B=1, C=32, h=16, w=24, D=8,  M=2, dot=True,  trans=2.50, rot=0.80, min_depth=0.5,  max_depth=10.0, seed=4),  # far out of view / behind camera
    "dot_ident":   dict(B=1, C=32, h=16, w=16, D=4,  M=1, dot=True,  trans=0.0,  rot=0.0,  min_depth=0.25, max_depth=20.0, seed=5),  # identity pose: the (w-1)/w shrink quirk
    "sad_rgb":     dict(B=2, C=3,  h=32, w=48, D=8,  M=1, dot=False, trans=0.10, rot=0.04, min_depth=0.5,  max_depth=50.0, seed=6),
    "sad_c32":     dict(B=1, C=32, h=16, w=24, D=8,  M=2, dot=False, trans=0.10, rot=0.04, min_depth=0.25, max_depth=20.0, seed=7),
}


def plane_sweep_inputs(synth, c):
    rng = np.random.RandomState(1000 + c["seed"])
    B, C, h, w, M = c["B"], c["C"], c["h"], c["w"], c["M"]
    image1 = synth.tensor("ps/%d/ref" % c["seed"], (B, C, h, w), seed=c["seed"], scale=4.0)
    image2s = [synth.tensor("ps/%d/meas%d" % (c["seed"], m), (B, C, h, w), seed=c["seed"], scale=4.0) for m in range(M)]
    pose1 = np.stack([_pose(rng, 0.5, 0.3) for _ in range(B)])
    pose2s = []
    for m in range(M):
        rel = np.stack([_pose(rng, c["trans"], c["rot"]) for _ in range(B)])
        pose2s.append(np.stack([pose1[b] @ rel[b] for b in range(B)]).astype(np.float32))
    K = np.stack([_K(0.9 * w * rng.uniform(0.9, 1.1), 0.9 * w * rng.uniform(0.9, 1.1), w / 2.0 + rng.uniform(-2, 2),
                     h / 2.0 + rng.uniform(-2, 2)) for _ in range(B)])
    return dict(image1=image1, image2s=image2s, pose1=pose1, pose2s=pose2s, K=K)


# ---------------------------------------------------------------------------------------------- hidden-state warp
HIDDEN_WARP_CASES = {
    "bottleneck":  dict(B=1, C=512, h=8, w=10, trans=0.10, rot=0.05, seed=11),   # fusionnet 320x256 bottleneck
    "batched":     dict(B=3, C=64,  h=8, w=8,  trans=0.30, rot=0.20, seed=12),
    "degenerate":  dict(B=2, C=32,  h=6, w=7,  trans=1.50, rot=0.90, seed=13),   # zero / tiny depths, points behind the camera
}


def hidden_warp_inputs(synth, c):
    rng = np.random.RandomState(2000 + c["seed"])
    B, C, h, w = c["B"], c["C"], c["h"], c["w"]
    image_src = synth.tensor("hw/%d/h" % c["seed"], (B, C, h, w), seed=c["seed"])
    depth = (0.3 + 4.0 * np.abs(synth.tensor("hw/%d/d" % c["seed"], (B, 1, h, w), seed=c["seed"]))).astype(np.float32)
    if c is HIDDEN_WARP_CASES["degenerate"] or c.get("seed") == 13:
        depth[:, :, 0, :] = 0.0
        depth[:, :, 1, 0:3] = 0.005
    trans = np.stack([_pose(rng, c["trans"], c["rot"]) for _ in range(B)])
    K = np.stack([_K(0.9 * w, 0.9 * w, w / 2.0, h / 2.0) for _ in range(B)])
    return dict(image_src=image_src, depth_dst=depth, trans=trans, K=K)


# ---------------------------------------------------------------------------------------------- depth re-projection
REPROJECT_CASES = {
    "small":  dict(B=1, H=64,  W=96,  trans=0.10, rot=0.05, seed=21),
    "c2":     dict(B=1, H=256, W=256, trans=0.12, rot=0.04, seed=22),
    "batch":  dict(B=2, H=64,  W=64,  trans=0.40, rot=0.30, seed=23),
}


def reproject_inputs(synth, c):
    rng = np.random.RandomState(3000 + c["seed"])
    B, H, W = c["B"], c["H"], c["W"]
    # smooth-ish depth with some structure so several source pixels land on one target pixel
    base = synth.smooth_image("rp/%d" % c["seed"], H, W, seed=c["seed"])[0:1]
    depth = (1.5 + 0.6 * base + 0.05 * synth.tensor("rp/%d/n" % c["seed"], (1, H, W), seed=c["seed"])).astype(np.float32)
    depth = np.clip(depth, 0.3, 20.0)
    previous_depth = np.stack([depth * (1.0 + 0.1 * b) for b in range(B)]).astype(np.float32)
    reference_pose = np.stack([_pose(rng, 0.5, 0.3) for _ in range(B)])
    rel = [_pose(rng, c["trans"], c["rot"]) for _ in range(B)]
    measurement_pose = np.stack([reference_pose[b] @ rel[b] for b in range(B)]).astype(np.float32)
    full_K = np.stack([_K(0.9 * W, 0.9 * W, W / 2.0 + 1.3, H / 2.0 - 0.7) for _ in range(B)])
    half_K = full_K.copy()
    half_K[:, 0:2, :] /= 2.0
    return dict(reference_pose=reference_pose, measurement_pose=measurement_pose, previous_depth=previous_depth,
                full_K=full_K, half_K=half_K)


# ---------------------------------------------------------------------------------------------- ConvLSTM cell
LSTM_CASES = {
    "warp":    dict(B=1, h=4, w=5, warp=True,  seed=31),
    "nowarp":  dict(B=2, h=4, w=4, warp=False, seed=32),
}


def lstm_inputs(synth, c):
    rng = np.random.RandomState(4000 + c["seed"])
    B, h, w = c["B"], c["h"], c["w"]
    weight = synth.make_state_dict({"lstm_cell.conv.weight": (2048, 1024, 3, 3)}, seed=c["seed"])["lstm_cell.conv.weight"]
    x = np.abs(synth.tensor("lstm/%d/x" % c["seed"], (B, 512, h, w), seed=c["seed"]))
    hh = synth.tensor("lstm/%d/h" % c["seed"], (B, 512, h, w), seed=c["seed"], scale=0.5)
    cc = synth.tensor("lstm/%d/c" % c["seed"], (B, 512, h, w), seed=c["seed"])
    depth = (0.5 + 3.0 * np.abs(synth.tensor("lstm/%d/d" % c["seed"], (B, 1, h, w), seed=c["seed"]))).astype(np.float32)
    depth[:, :, 0, 0] = 0.0          # exercises the <= 0.01 mask (convlstm.py:32,40-41)
    previous_pose = np.stack([_pose(rng, 0.5, 0.3) for _ in range(B)])
    rel = [_pose(rng, 0.1, 0.05) for _ in range(B)]
    current_pose = np.stack([previous_pose[b] @ rel[b] for b in range(B)]).astype(np.float32)
    K = np.stack([_K(0.9 * w, 0.9 * w, w / 2.0, h / 2.0) for _ in range(B)])
    return dict(weight=weight, x=x, h=hh, c=cc, depth=depth, previous_pose=previous_pose, current_pose=current_pose, K=K)


# ---------------------------------------------------------------------------------------------- whole modules
MODULE_CASE = dict(B=1, H=64, W=96, D=64, seed=41)


def module_inputs(synth, c):
    B, H, W, D = c["B"], c["H"], c["W"], c["D"]
    image = np.stack([synth.smooth_image("mod/%d/img%d" % (c["seed"], b), H, W, seed=c["seed"]) for b in range(B)])
    cost_volume = synth.tensor("mod/%d/cv" % c["seed"], (B, D, H // 2, W // 2), seed=c["seed"], scale=5.0)
    depth_est = (0.5 + 2.0 * np.abs(synth.tensor("mod/%d/de" % c["seed"], (B, 1, H // 32, W // 32), seed=c["seed"]))).astype(np.float32)
    depth_est[:, :, 0, 0] = 0.0
    K = synth.intrinsics(H, W)
    lstm_K = K.copy()
    lstm_K[0:2, :] /= 32.0
    pose0 = np.stack([synth.camera_pose(3)] * B)
    pose1 = np.stack([synth.camera_pose(4)] * B)
    return dict(image=image, cost_volume=cost_volume, depth_est=depth_est, lstm_K=np.stack([lstm_K] * B), pose0=pose0, pose1=pose1)


# ---------------------------------------------------------------------------------------------- training (row f3)
# Gradient cases: which forward cases get a seeded upstream gradient pushed through the REFERENCE under autograd
# (oracle/make_golden_training.py -> tests/golden/training.npz).
SWEEP_GRAD_CASES = ["dot_small", "dot_c1", "dot_m3", "dot_wide"]
HIDDEN_WARP_GRAD_CASES = ["bottleneck", "batched", "degenerate"]
LSTM_GRAD_CASES = ["warp", "nowarp"]


def upstream(synth, key, shape, seed):
    """Seeded upstream gradient d loss / d output for a gradient case."""
    return synth.tensor("grad/" + key, shape, seed=seed)


LOSS_CASE = dict(B=2, H=64, W=96, seed=51, weights=[1.0, 0.5, 2.0, 1.0, 1.5])
LOSS_TYPES = ["L1", "L1-inv", "L1-rel", "Huber"]


def loss_inputs(synth, c):
    """Ground-truth depth with invalid (zero) pixels and five prediction scales (1/16 ... full), as
    update_losses receives them (fusionnet/run-training.py:270)."""
    B, H, W = c["B"], c["H"], c["W"]
    gt = (0.5 + 3.0 * np.abs(synth.tensor("loss/%d/gt" % c["seed"], (B, H, W), seed=c["seed"]))).astype(np.float32)
    gt[:, 0:5, :] = 0.0
    gt[:, :, 7::13] = 0.0
    preds = []
    for k, s in enumerate((16, 8, 4, 2, 1)):
        p = (0.4 + 3.0 * np.abs(synth.tensor("loss/%d/p%d" % (c["seed"], k), (B, H // s, W // s), seed=c["seed"]))).astype(np.float32)
        preds.append(p)
    preds[4][:, 10:20, 10:20] = gt[:, 10:20, 10:20]          # exact hits: sign(0) = 0 in every loss
    preds[4][:, 30:34, 30:34] = gt[:, 30:34, 30:34] + 0.25   # inside the quadratic zone of smooth_l1
    return dict(groundtruth=gt, predictions=preds)


# ---------------------------------------------------------------------------------------------- whole training step
# BASELINE.json config 5 scaled down: batch 2, subsequence of 3 (two reference / measurement pairs), 128 x 160, 64 planes
# (large enough that the train-mode BatchNorms at 1/32 resolution see 40 samples: at 64 x 96 the gradients of the reference
# differ from THEMSELVES by 1e-3 between 1 and 8 threads).
TRAIN_STEP_CASE = dict(B=2, T=3, H=128, W=160, D=64, seed=61, loss_type="L1-inv")
# (module, parameter, slice or None): gradients stored in full
TRAIN_STEP_FULL_GRADS = [
    ("fe", "layer1.0.weight", None), ("fe", "layer5.0.0.layers.6.weight", None), ("fpn", "fpn.inner_blocks.4.weight", None),
    ("fpn", "fpn.layer_blocks.0.bias", None), ("cve", "aggregator0.0.weight", None), ("cve", "encoder_block3.standard_convolution.conv2.1.weight", None),
    ("lstm", "lstm_cell.conv.weight", (slice(0, 16), slice(0, 16))), ("cvd", "refine.1.0.weight", None), ("cvd", "depth_layer_full.0.bias", None),
    ("cvd", "decoder_block1.up_convolution.conv.0.weight", (slice(0, 8), slice(0, 8))),
]


def train_step_inputs(synth, c):
    B, T_, H, W = c["B"], c["T"], c["H"], c["W"]
    images, depths, poses = [], [], []
    for t in range(T_):
        images.append(np.stack([synth.smooth_image("train/%d/img%d_%d" % (c["seed"], b, t), H, W, seed=c["seed"] + b) for b in range(B)]))
        d = (0.6 + 2.5 * np.abs(synth.tensor("train/%d/depth%d" % (c["seed"], t), (B, H, W), seed=c["seed"]))).astype(np.float32)
        d[:, 0:3, :] = 0.0                 # invalid ground truth: masked in the loss and in the hidden-state warp
        d[:, :, 5::17] = 0.0
        depths.append(d)
        poses.append(np.stack([synth.camera_pose(t + b) for b in range(B)]))
    K = np.stack([synth.intrinsics(H, W)] * B)
    return dict(images=images, depths=depths, poses=poses, K=K)
