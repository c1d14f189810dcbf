"""Shared helpers for the parity tests (oracle on CPU vs the CUDA path through the reference-facing modules)."""
import numpy as np
import torch

T = torch.from_numpy


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def oracle_weights(oracle, synth, seed, n_depth_levels=64):
    shapes = oracle.state_dict_shapes(n_depth_levels)
    return {tag: {k: T(v) for k, v in synth.make_state_dict(shapes[tag], seed=seed).items()} for tag in shapes}


def build_product_modules(weights, device="cuda", n_depth_levels=64, pairnet=False):
    from dvmvs.pipeline import build_modules
    return build_modules(weights, device=device, n_depth_levels=n_depth_levels, pairnet=pairnet)


class ProductState:
    def __init__(self):
        self.lstm_state = None
        self.previous_depth = None
        self.previous_pose = None


def product_fusionnet_step(mods, state, reference_image, reference_pose, measurement_images, measurement_poses, full_K,
                           min_depth=0.25, max_depth=20.0, n_depth_levels=64):
    """The loop body of the reference's fusionnet/run-testing.py:145-202, spelled out the way that script spells it
    (including F.interpolate for the 1/16 nearest down-sampling), calling the drop-in modules (all tensors CUDA)."""
    from dvmvs.utils import (cost_volume_fusion, get_non_differentiable_rectangle_depth_estimation,
                             get_warp_grid_for_cost_volume_calculation)
    B, _, H, W = reference_image.shape
    device = reference_image.device
    half_K = full_K.clone()
    half_K[:, 0:2, :] = half_K[:, 0:2, :] / 2.0
    lstm_K = full_K.clone()
    lstm_K[:, 0:2, :] = lstm_K[:, 0:2, :] / 32.0
    warp_grid = get_warp_grid_for_cost_volume_calculation(width=W // 2, height=H // 2, device=device)
    meas_half = []
    for im in measurement_images:
        half, _, _, _ = mods["fpn"](*mods["fe"](im))
        meas_half.append(half)
    f2, f4, f8, f16 = mods["fpn"](*mods["fe"](reference_image))
    cv = cost_volume_fusion(image1=f2, image2s=meas_half, pose1=reference_pose, pose2s=measurement_poses, K=half_K,
                            warp_grid=warp_grid, min_depth=min_depth, max_depth=max_depth, n_depth_levels=n_depth_levels,
                            device=device, dot_product=True)
    s0, s1, s2, s3, bottom = mods["cve"](features_half=f2, features_quarter=f4, features_one_eight=f8,
                                         features_one_sixteen=f16, cost_volume=cv)
    if "lstm" in mods:
        if state.previous_depth is not None:
            de = get_non_differentiable_rectangle_depth_estimation(reference_pose_torch=reference_pose,
                                                                   measurement_pose_torch=state.previous_pose,
                                                                   previous_depth_torch=state.previous_depth,
                                                                   full_K_torch=full_K, half_K_torch=half_K,
                                                                   original_height=H, original_width=W)
            de = torch.nn.functional.interpolate(input=de, scale_factor=(1.0 / 16.0), mode="nearest")
        else:
            de = torch.zeros(size=(B, 1, H // 32, W // 32), device=device)
        state.lstm_state = mods["lstm"](current_encoding=bottom, current_state=state.lstm_state,
                                        previous_pose=state.previous_pose, current_pose=reference_pose,
                                        estimated_current_depth=de, camera_matrix=lstm_K)
        bottom = state.lstm_state[0]
    pred = mods["cvd"](reference_image, s0, s1, s2, s3, bottom)[0]
    state.previous_depth = pred.view(B, 1, H, W)
    state.previous_pose = reference_pose
    return pred, state
