"""Deterministic synthetic weights / posed RGB streams (numpy only, no torch, no package imports).

Shared by the product benchmark (bench.py), the tests and the golden-vector generator
(oracle/make_golden.py loads this file by path inside the real-reference process), so that the
oracle and the CUDA path always see identical tensors.  Follows SURVEY.md section 8(d): reference
pose = identity, measurement k = 0.1*k m along x (+0.05*k along y), yaw 0.02*k rad, clips advance
0.1 m per keyframe, K: fx = fy = 0.9*W, cx = W/2, cy = H/2, depth range 0.25..20 m.

numpy.random.RandomState streams are frozen across numpy versions, so (key, shape, seed) -> tensor is
reproducible here and on the GPU box.
"""
import zlib

import numpy as np

MEAN_RGB = (0.485, 0.456, 0.406)   # reference dvmvs/fusionnet/run-testing.py:53-55
STD_RGB = (0.229, 0.224, 0.225)
SCALE_RGB = 255.0


def _rng(key, seed):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def make_state_dict(shapes, seed=0):
    """shapes: {state_dict key: tuple shape} (taken from module.state_dict()).  Returns
    {key: np.ndarray} with He-initialised conv weights and O(1) BatchNorm statistics so that
    activations stay well-scaled through ~100 layers (BN folded or not)."""
    keys = list(shapes.keys())
    out = {}
    for key in keys:
        shape = tuple(int(s) for s in shapes[key])
        r = _rng(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[key] = np.zeros(shape, dtype=np.int64)
        elif leaf == "running_mean":
            out[key] = (0.1 * r.randn(*shape)).astype(np.float32)
        elif leaf == "running_var":
            out[key] = r.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "weight" and len(shape) == 1:        # BatchNorm gamma
            out[key] = r.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "bias":
            out[key] = (0.1 * r.randn(*shape)).astype(np.float32)
        elif leaf == "weight" and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            if "depth_layer" in key:          # keep the sigmoid depth heads out of saturation
                gain = 0.1
            elif ".layers.6." in key:         # MnasNet linear bottleneck (no ReLU, residual adds follow)
                gain = 0.5
            elif key.startswith("fpn.") or "lstm_cell" in key or key.startswith("layer1.6."):   # no ReLU after
                gain = 0.7
            else:                             # followed by BN + ReLU: He
                gain = 1.4142135
            out[key] = (gain * np.sqrt(1.0 / fan_in) * r.randn(*shape)).astype(np.float32)
        else:
            raise KeyError("synth_data.make_state_dict: unexpected key %r shape %r" % (key, shape))
    return out


def tensor(key, shape, seed=0, scale=1.0, positive=False):
    """Seeded N(0, scale) tensor (|.| if positive) for op-level inputs."""
    x = scale * _rng(key, seed).randn(*[int(s) for s in shape])
    if positive:
        x = np.abs(x)
    return x.astype(np.float32)


def smooth_image(key, H, W, seed=0):
    """Normalised RGB frame (3,H,W) fp32: bilinearly-upsampled low-frequency noise plus fine noise in
    the uint8 range, normalised with the reference's mean/std (run-testing.py:53-55)."""
    r = _rng(key, seed)
    gh, gw = max(2, H // 16), max(2, W // 16)
    coarse = r.uniform(0.0, 255.0, size=(3, gh, gw))
    ys = np.linspace(0, gh - 1, H)
    xs = np.linspace(0, gw - 1, W)
    y0 = np.minimum(np.floor(ys).astype(int), gh - 2)
    x0 = np.minimum(np.floor(xs).astype(int), gw - 2)
    fy = (ys - y0)[None, :, None]
    fx = (xs - x0)[None, None, :]
    c00 = coarse[:, y0][:, :, x0]
    c01 = coarse[:, y0][:, :, x0 + 1]
    c10 = coarse[:, y0 + 1][:, :, x0]
    c11 = coarse[:, y0 + 1][:, :, x0 + 1]
    img = (c00 * (1 - fy) * (1 - fx) + c01 * (1 - fy) * fx + c10 * fy * (1 - fx) + c11 * fy * fx)
    img = np.clip(img + r.uniform(-20.0, 20.0, size=img.shape), 0.0, 255.0)
    img = np.floor(img) / SCALE_RGB
    for c in range(3):
        img[c] = (img[c] - MEAN_RGB[c]) / STD_RGB[c]
    return img.astype(np.float32)


def intrinsics(H, W):
    return np.array([[0.9 * W, 0.0, W / 2.0], [0.0, 0.9 * W, H / 2.0], [0.0, 0.0, 1.0]], dtype=np.float32)


def _yaw(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def camera_pose(step):
    """cam-to-world pose after `step` units of 0.1 m motion (x, half as much y) and 0.02 rad yaw each."""
    P = np.eye(4)
    P[:3, :3] = _yaw(0.02 * step)
    P[:3, 3] = [0.1 * step, 0.05 * step, 0.0]
    return P.astype(np.float32)


def make_clip(clip_seed, n_frames, H, W, n_meas):
    """One synthetic posed RGB stream.  Returns dict with
    images (n_frames + n_meas, 3, H, W) fp32, poses (n_frames + n_meas, 4, 4) fp32, K (3,3) fp32 and
    frames: list of (reference_index, [measurement indices]); keyframe t uses the n_meas previous
    stream frames as measurement frames (pose distance 0.1*k m, matching Config.test_keyframe_pose_distance)."""
    total = n_frames + n_meas
    images = np.stack([smooth_image("clip%d/frame%d" % (clip_seed, i), H, W, seed=clip_seed) for i in range(total)])
    poses = np.stack([camera_pose(i) for i in range(total)])
    frames = [(t + n_meas, [t + n_meas - k for k in range(1, n_meas + 1)]) for t in range(n_frames)]
    return {"images": images, "poses": poses, "K": intrinsics(H, W), "frames": frames}
