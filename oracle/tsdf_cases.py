"""Seeded TSDF-integration cases shared by the golden generator (reference side) and the tests (numpy only).
Each case: a small voxel volume, a few posed RGB-D frames (depth with holes = 0).  Dtypes follow the reference's caller
(run-tsdf-reconstruction.py:500-560: uint8 RGB images, float32 predicted depth maps, float64 poses); the "behind" case feeds
float64 depth and float32 colours instead -- the CPU path of the reference accepts both and computes differently in them."""
import numpy as np

CASES = {
    #  name: volume bounds (3,2), voxel size, image (h, w), frames, seed
    "small":   dict(bounds=[[-0.6, 0.6], [-0.5, 0.5], [0.2, 1.6]], voxel=0.04, hw=(24, 32), frames=3, seed=1),
    "ragged":  dict(bounds=[[-0.33, 0.41], [-0.27, 0.30], [0.10, 1.05]], voxel=0.035, hw=(30, 21), frames=4, seed=2),   # dims not multiples of anything
    "behind":  dict(bounds=[[-0.5, 0.5], [-0.5, 0.5], [-0.6, 0.8]], voxel=0.05, hw=(16, 20), frames=2, seed=3),        # voxels behind the camera / outside the frustum
}


def _pose(rng, t_scale, angle):
    ax = rng.randn(3)
    ax /= np.linalg.norm(ax)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(angle) * Kx + (1 - np.cos(angle)) * Kx @ Kx
    P = np.eye(4)
    P[:3, :3] = R
    P[:3, 3] = t_scale * rng.uniform(-1, 1, size=3)
    return P                                             # float64, like poses parsed from poses.txt


def inputs(case):
    c = CASES[case]
    rng = np.random.RandomState(500 + c["seed"])
    h, w = c["hw"]
    K = np.array([[0.9 * w, 0, w / 2.0 - 0.3], [0, 0.92 * w, h / 2.0 + 0.2], [0, 0, 1.0]])     # float64; integrate() casts to float32
    frames = []
    for i in range(c["frames"]):
        depth = (0.5 + 0.6 * rng.rand(h, w) + 0.2 * np.sin(np.arange(w) / 3.0)[None, :]).astype(np.float64)
        depth[rng.rand(h, w) < 0.08] = 0.0                # invalid pixels
        color = np.floor(rng.uniform(0, 256, size=(h, w, 3))).astype(np.uint8)                # RGB
        if case == "behind":
            color = color.astype(np.float32)
        else:
            depth = depth.astype(np.float32)
        frames.append(dict(color=color, depth=depth, pose=_pose(rng, 0.08, 0.06 * (i + 1)), weight=1.0 if i != 1 else (2.0 if case != "ragged" else 0.3)))
    return dict(bounds=np.array(c["bounds"], dtype=np.float64), voxel=c["voxel"], K=K, frames=frames)
