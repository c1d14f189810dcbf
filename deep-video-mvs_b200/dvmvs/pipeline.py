"""The per-keyframe call sequence of the reference's test drivers (fusionnet/run-testing.py:145-202,
pairnet/run-testing.py:139-164) as a function over the drop-in modules, for callers that do not want to copy
the script's loop body (bench.py, smoke test, examples).  It makes exactly the module / utils calls the script
makes, with the same keyword arguments."""
import torch

from .utils import (cost_volume_fusion, get_non_differentiable_rectangle_depth_estimation,
                    get_warp_grid_for_cost_volume_calculation)


class KeyframeState:
    """Recurrent state the caller carries between keyframes of one clip (run-testing.py:86-88)."""

    def __init__(self):
        self.lstm_state = None
        self.previous_depth = None
        self.previous_pose = None

    def reset(self):          # "TRACKING LOST" (run-testing.py:97-101)
        self.__init__()


def build_modules(weights, device="cuda", n_depth_levels=64, pairnet=False):
    """Constructs the modules, strict-loads `weights` (dict 'fe','fpn','cve'[,'lstm'],'cvd' -> state dict), eval()."""
    from .config import Config
    old = Config.train_n_depth_levels
    Config.train_n_depth_levels = n_depth_levels          # aggregator0 has D+32 inputs (fusionnet/model.py:170)
    try:
        if pairnet:
            from .pairnet import model as mm
        else:
            from .fusionnet import model as mm
        mods = {"fe": mm.FeatureExtractor(), "fpn": mm.FeatureShrinker(), "cve": mm.CostVolumeEncoder(), "cvd": mm.CostVolumeDecoder()}
        if not pairnet:
            mods["lstm"] = mm.LSTMFusion()
    finally:
        Config.train_n_depth_levels = old
    for tag, m in mods.items():
        m.load_state_dict(weights[tag], strict=True)
        m.to(device).eval()
    return mods


def keyframe(mods, state, reference_image, reference_pose, measurement_images, measurement_poses, full_K,
             min_depth=0.25, max_depth=20.0, n_depth_levels=64):
    """One keyframe for B independent clips (tensors batched on dim 0, all CUDA).  With 'lstm' in mods this is the
    fusionnet loop body, without it the pairnet one.  Returns (depth (B,H,W), state)."""
    B, _, H, W = reference_image.shape
    device = reference_image.device
    half_K = full_K.clone()
    half_K[:, 0:2, :] = half_K[:, 0:2, :] / 2.0
    warp_grid = None    # ignored by the fused kernel (kept in the signature for API compatibility)
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
        lstm_K = full_K.clone()
        lstm_K[:, 0:2, :] = lstm_K[:, 0:2, :] / 32.0
        if state.previous_depth is not None:
            de = get_non_differentiable_rectangle_depth_estimation(reference_pose_torch=reference_pose,
                                                                   measurement_pose_torch=state.previous_pose,
                                                                   previous_depth_torch=state.previous_depth,
                                                                   full_K_torch=full_K, half_K_torch=half_K,
                                                                   original_height=H, original_width=W)
            de = de[:, :, ::16, ::16].contiguous()      # == F.interpolate(scale_factor=1/16, mode='nearest') (run-testing.py:187-189)
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
