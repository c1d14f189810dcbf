"""SURVEY section 8 row f4 -- TSDF integration (reference: sample-data/run-tsdf-reconstruction.py TSDFVolume.integrate).
CPU: the numpy oracle against goldens of the UNMODIFIED reference script (oracle/make_golden_tsdf.py), bit for bit.
GPU: dvmvs.tsdf.TSDFVolume (one launch of dvmvs_tsdf_integrate per frame, through the C-ABI) against the same goldens and, at a
production-sized volume, against the oracle -- integer / fp32 state compared with array_equal, no tolerance."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "deep-video-mvs_b200"))
import tsdf_cases  # noqa: E402
import tsdf_oracle  # noqa: E402

GOLD = np.load(os.path.join(REPO, "tests", "golden", "tsdf.npz"))


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("case", sorted(tsdf_cases.CASES))
def test_oracle_equals_reference_goldens_bit_for_bit(case):
    inp = tsdf_cases.inputs(case)
    vol = tsdf_oracle.TSDFVolume(inp["bounds"], inp["voxel"])
    assert np.array_equal(vol.vol_dim, GOLD[case + "/vol_dim"]) and np.array_equal(vol.vol_origin, GOLD[case + "/vol_origin"])
    for i, fr in enumerate(inp["frames"]):
        assert vol.integrate(fr["color"], fr["depth"], inp["K"], fr["pose"], fr["weight"]) > 0
        assert _same(vol.tsdf, GOLD["%s/tsdf_after_%d" % (case, i)]), (case, i)
        assert _same(vol.weight, GOLD["%s/weight_after_%d" % (case, i)]), (case, i)
        assert _same(vol.color, GOLD["%s/color_after_%d" % (case, i)]), (case, i)


@pytest.mark.parametrize("case", sorted(tsdf_cases.CASES))
def test_frustum_bounds_host_logic_equals_reference(case):
    from dvmvs.tsdf import TSDFFusion
    inp = tsdf_cases.inputs(case)
    got = TSDFFusion.calculate_volume_bounds([f["depth"] for f in inp["frames"]], [f["pose"] for f in inp["frames"]], inp["K"])
    assert np.array_equal(got, GOLD[case + "/frustum_bounds"])


def test_volume_refuses_cpu_mode():
    from dvmvs.tsdf import TSDFVolume
    with pytest.raises(RuntimeError):
        TSDFVolume(np.array([[0, 1], [0, 1], [0, 1.0]]), 0.1, use_gpu=False)
    with pytest.raises(AssertionError):
        TSDFVolume(np.zeros((2, 3)), 0.1)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(tsdf_cases.CASES))
def test_gpu_volume_equals_reference_goldens_bit_for_bit(case):
    from dvmvs.tsdf import TSDFVolume
    inp = tsdf_cases.inputs(case)
    vol = TSDFVolume(inp["bounds"], inp["voxel"])
    assert np.array_equal(vol._vol_dim, GOLD[case + "/vol_dim"]) and np.array_equal(vol._vol_origin, GOLD[case + "/vol_origin"])
    for i, fr in enumerate(inp["frames"]):
        vol.integrate(fr["color"], fr["depth"], inp["K"], fr["pose"], obs_weight=fr["weight"])
        tsdf, color = vol.get_volume()
        weight = vol.get_volume_tensors()[1].cpu().numpy()
        for name, got in (("tsdf", tsdf), ("weight", weight), ("color", color)):
            want = GOLD["%s/%s_after_%d" % (case, name, i)]
            assert _same(got, want), "%s frame %d: %s differs in %d voxels (max %g)" % (
                case, i, name, int((got != want).sum()), float(np.nanmax(np.abs(got - want))))


@pytest.mark.gpu
def test_gpu_volume_accepts_device_tensors_and_counts_updates():
    import torch
    from dvmvs.tsdf import TSDFVolume
    inp = tsdf_cases.inputs("small")
    a, b = TSDFVolume(inp["bounds"], inp["voxel"]), TSDFVolume(inp["bounds"], inp["voxel"])
    orc = tsdf_oracle.TSDFVolume(inp["bounds"], inp["voxel"])
    n = 0
    for fr in inp["frames"]:
        a.integrate(fr["color"], fr["depth"], inp["K"], fr["pose"], obs_weight=fr["weight"])
        b.integrate(torch.from_numpy(fr["color"]).cuda(), torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(inp["K"]),
                    torch.from_numpy(fr["pose"]), obs_weight=fr["weight"])
        n += orc.integrate(fr["color"], fr["depth"], inp["K"], fr["pose"], fr["weight"])
    for x, y in zip(a.get_volume_tensors(), b.get_volume_tensors()):
        assert torch.equal(x, y)
    assert a.updated_voxels() == b.updated_voxels() == n


@pytest.mark.gpu
def test_gpu_volume_at_production_size_equals_oracle():
    """A 3.9 M-voxel volume (4 cm voxels over a 8 x 6.4 x 4.8 m room, the reference's default voxel size), 320 x 256 frames."""
    from dvmvs.tsdf import TSDFVolume
    rng = np.random.RandomState(11)
    h, w = 256, 320
    K = np.array([[250.0, 0, 160.3], [0, 251.0, 127.6], [0, 0, 1]])
    bounds = np.array([[-4.0, 4.0], [-3.2, 3.2], [0.0, 4.8]])
    vol, orc = TSDFVolume(bounds, 0.04), tsdf_oracle.TSDFVolume(bounds, 0.04)
    assert int(np.prod(vol._vol_dim)) == 200 * 160 * 120
    total = 0
    for i in range(3):
        yy, xx = np.mgrid[0:h, 0:w]
        depth = (2.0 + 0.8 * np.sin(xx / 40.0 + i) * np.cos(yy / 30.0) + 0.01 * rng.rand(h, w)).astype(np.float32)
        depth[rng.rand(h, w) < 0.05] = 0
        color = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        pose = np.eye(4)
        pose[:3, 3] = [0.1 * i, -0.05 * i, 0.02 * i]
        c, s = np.cos(0.05 * i), np.sin(0.05 * i)
        pose[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        vol.integrate(color, depth, K, pose, obs_weight=1.0)
        total += orc.integrate(color, depth, K, pose, 1.0)
    tsdf, color_vol = vol.get_volume()
    assert _same(tsdf, orc.tsdf) and _same(color_vol, orc.color)
    assert _same(vol.get_volume_tensors()[1].cpu().numpy(), orc.weight)
    assert vol.updated_voxels() == total and total > 100000


def _rot(axis, angle):
    axis = np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * Kx + (1 - np.cos(angle)) * Kx @ Kx


@pytest.mark.gpu
@pytest.mark.parametrize("offset", [0.0, 150.0])
def test_gpu_frustum_cull_never_drops_a_voxel_the_reference_updates(offset):
    """The kernel's float32 pre-test must be conservative: cameras INSIDE the volume, looking along every axis and obliquely
    (many voxels on the image border and next to the camera plane), volume far from the origin (offset: float32 world
    coordinates with 1e-5 m resolution), noisy depth.  Whole volumes equal the oracle's."""
    from dvmvs.tsdf import TSDFVolume
    rng = np.random.RandomState(23)
    h, w = 96, 128
    K = np.array([[100.0, 0, 63.5], [0, 100.0, 47.5], [0, 0, 1]])
    bounds = np.array([[-2.0, 2.0], [-1.6, 1.6], [-2.0, 2.0]]) + offset
    vol, orc = TSDFVolume(bounds, 0.04), tsdf_oracle.TSDFVolume(bounds, 0.04)
    views = [([0, 1, 0], 0.0), ([0, 1, 0], np.pi / 2), ([0, 1, 0], np.pi), ([1, 0, 0], np.pi / 2), ([1, 0, 0], -np.pi / 2),
             ([1, 1, 0], 0.7), ([1, 2, 3], 2.1), ([0, 0, 1], np.pi / 4)]
    total = 0
    for i, (axis, angle) in enumerate(views):
        depth = (0.3 + 1.5 * rng.rand(h, w)).astype(np.float32)
        color = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        pose = np.eye(4)
        pose[:3, :3] = _rot(axis, angle)
        pose[:3, 3] = offset + rng.uniform(-0.5, 0.5, size=3)
        vol.integrate(color, depth, K, pose, obs_weight=1.0 + i)
        total += orc.integrate(color, depth, K, pose, 1.0 + i)
    tsdf, color_vol = vol.get_volume()
    assert _same(vol.get_volume_tensors()[1].cpu().numpy(), orc.weight)
    assert _same(tsdf, orc.tsdf) and _same(color_vol, orc.color)
    assert vol.updated_voxels() == total and total > 50000
