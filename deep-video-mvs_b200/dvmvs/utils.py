"""Free functions of the reference's dvmvs/utils.py that the inference scripts import (run-testing.py:7-8,
convlstm.py:4), same names / argument order / error behaviour, backed by the sm_100a kernels.

Hot:   cost_volume_fusion, calculate_cost_volume_by_warping, warp_frame_depth,
       get_non_differentiable_rectangle_depth_estimation, get_warp_grid_for_cost_volume_calculation
Glue:  pose_distance, is_pose_available, InferenceTimer, save_results, save_predictions, visualize_predictions
"""
import os

import numpy as np
import torch

from . import _ops as ops
from .errors import compute_errors


# ------------------------------------------------------------------------------------------------ geometry (hot)
def get_warp_grid_for_cost_volume_calculation(width, height, device):
    """Reference utils.py:34-42: homogeneous pixel grid (3, h*w).  Kept for API compatibility -- the fused kernel derives
    pixel coordinates from its thread index and ignores the grid it is handed."""
    ys, xs = torch.meshgrid(torch.arange(int(height), dtype=torch.float32), torch.arange(int(width), dtype=torch.float32),
                            indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(int(height) * int(width))], dim=0).to(device)


def _check_sweep_args(image1, image2s, pose2s):
    ops.require_cuda_f32(image1, "image1")
    if image1.dim() != 4:
        raise ValueError("image1 must have shape (B, C, H, W), got %s" % (tuple(image1.shape),))
    if len(image2s) != len(pose2s) or len(image2s) == 0:
        raise ValueError("need as many measurement poses as measurement images (>= 1)")


def cost_volume_fusion(image1, image2s, pose1, pose2s, K, warp_grid, min_depth, max_depth, n_depth_levels, device, dot_product):
    """Reference utils.py:89-107.  ONE fused launch over all planes and all measurement frames; the warped feature
    tensors are never materialised.  Returns (B, D, h, w) fp32 (channels_last strides)."""
    image2s, pose2s = list(image2s), list(pose2s)
    _check_sweep_args(image1, image2s, pose2s)
    if not (torch.is_grad_enabled() and any(t.requires_grad for t in [image1] + image2s)) and all(
            getattr(t, "_dvmvs_act", None) is None for t in [image1] + image2s):
        # foreign tensors (the script path): split kernels + sweep replayed as one CUDA graph (_base.graphed_call); tensors that
        # carry their producer's fp16 planes go straight to the kernel (the engines capture that themselves)
        from ._base import graphed_call
        return graphed_call(("cost_volume_fusion",), _cost_volume_fusion, (image1, image2s, pose1, pose2s, K, float(min_depth), float(max_depth),
                                                                            int(n_depth_levels), bool(dot_product)), {})
    return _cost_volume_fusion(image1, image2s, pose1, pose2s, K, min_depth, max_depth, n_depth_levels, dot_product)


def _cost_volume_fusion(image1, image2s, pose1, pose2s, K, min_depth, max_depth, n_depth_levels, dot_product):
    device, warp_grid = image1.device, None
    if torch.is_grad_enabled() and any(t.requires_grad for t in [image1] + image2s):
        # training (run-training.py:231): same forward kernel, hand-written backward kernel (dvmvs/training.py, row f3)
        if not dot_product:
            raise NotImplementedError("only the dot-product cost volume is differentiable (the SAD branch is used by the baselines' "
                                      "inference only, utils.py:83-84)")
        from .training import plane_sweep_cost_volume
        return plane_sweep_cost_volume(image1, image2s, pose1, pose2s, K, min_depth, max_depth, n_depth_levels)
    if ops.sweep_uses_tc(bool(dot_product), image1.shape[1], int(n_depth_levels), len(image2s)):
        # tensor-core form (csrc/sweep_tc.cu): correlate the epipolar band on tcgen05, blend four scalars per sample.  Reads the
        # fp16 (hi, lo) planes the FPN's output convolution emitted (a split kernel stages them for foreign tensors).
        ref_pair = ops.act_pair(ops.to_act(image1, "image1"))
        meas_pairs = [ops.act_pair(ops.to_act(t, "image2")) for t in image2s]
        cost = ops.plane_sweep_tc(ref_pair, meas_pairs, pose1, pose2s, K, min_depth, max_depth, n_depth_levels, terms=ops.sweep_terms())
        return ops.to_api(cost)
    ref = ops.to_nhwc(image1, "image1")
    if ops.SWEEP_FP16 and dot_product and ref.shape[-1] == 32 and ref.shape[2] >= 2:
        # experimental (DVMVS_SWEEP_FP16=1): gather the fp16 "hi" plane of the measurement features -- already there when they
        # come out of a tensor-core convolution, split off once otherwise
        hi = [ops.to_act(t, "image2").get_planes()[0] for t in image2s]
        return ops.to_api(ops.plane_sweep_h16(ref, hi, pose1, pose2s, K, min_depth, max_depth, n_depth_levels))
    meas = [ops.to_nhwc(t, "image2") for t in image2s]
    cost = ops.plane_sweep(ref, meas, pose1, pose2s, K, min_depth, max_depth, n_depth_levels, bool(dot_product))
    return ops.to_api(cost)


def calculate_cost_volume_by_warping(image1, image2, pose1, pose2, K, warp_grid, min_depth, max_depth, n_depth_levels, device,
                                     dot_product):
    """Reference utils.py:45-86 (single measurement frame; the training script's entry point)."""
    return cost_volume_fusion(image1, [image2], pose1, [pose2], K, warp_grid, min_depth, max_depth, n_depth_levels, device,
                              dot_product)


def get_non_differentiable_rectangle_depth_estimation(reference_pose_torch, measurement_pose_torch, previous_depth_torch,
                                                      full_K_torch, half_K_torch, original_width, original_height):
    """Reference utils.py:110-154.  One scatter kernel (atomicMax on z >= 0 as uint == 'first in z-descending order
    wins'); no argsort, no device->host round trip (the reference syncs at utils.py:148)."""
    B = reference_pose_torch.shape[0]
    H, W = int(original_height), int(original_width)
    if previous_depth_torch.numel() != B * H * W:
        raise ValueError("previous_depth_torch must hold B*H*W = %d values, got shape %s" % (B * H * W, tuple(previous_depth_torch.shape)))
    from ._base import graphed_call
    return graphed_call(("depth_reproject", H, W), lambda a, b, c, d, e: ops.depth_reproject(a, b, c, d, e, H, W),
                        (reference_pose_torch, measurement_pose_torch, previous_depth_torch, full_K_torch, half_K_torch), {})


def warp_frame_depth(image_src, depth_dst, src_trans_dst, camera_matrix, normalize_points=False, sampling_mode='bilinear'):
    """Reference utils.py:205-258 (kornia's warp_frame_depth).  Same argument checks / exception types."""
    if not isinstance(image_src, torch.Tensor):
        raise TypeError(f"Input image_src type is not a torch.Tensor. Got {type(image_src)}.")
    if not len(image_src.shape) == 4:
        raise ValueError(f"Input image_src musth have a shape (B, D, H, W). Got: {image_src.shape}")
    if not isinstance(depth_dst, torch.Tensor):
        raise TypeError(f"Input depht_dst type is not a torch.Tensor. Got {type(depth_dst)}.")
    if not (len(depth_dst.shape) == 4 and depth_dst.shape[-3] == 1):
        raise ValueError(f"Input depth_dst musth have a shape (B, 1, H, W). Got: {depth_dst.shape}")
    if not isinstance(src_trans_dst, torch.Tensor):
        raise TypeError(f"Input src_trans_dst type is not a torch.Tensor. Got {type(src_trans_dst)}.")
    if not (len(src_trans_dst.shape) == 3 and tuple(src_trans_dst.shape[-2:]) == (4, 4)):
        raise ValueError(f"Input src_trans_dst must have a shape (B, 4, 4). Got: {src_trans_dst.shape}.")
    if not isinstance(camera_matrix, torch.Tensor):
        raise TypeError(f"Input camera_matrix type is not a torch.Tensor. Got {type(camera_matrix)}.")
    if not (len(camera_matrix.shape) == 3 and tuple(camera_matrix.shape[-2:]) == (3, 3)):
        raise ValueError(f"Input camera_matrix must have a shape (B, 3, 3). Got: {camera_matrix.shape}.")
    if normalize_points or sampling_mode != 'bilinear':
        raise NotImplementedError("the sm_100a kernel implements normalize_points=False, sampling_mode='bilinear' "
                                  "(the only combination the reference uses, convlstm.py:33-38)")
    if torch.is_grad_enabled() and image_src.requires_grad:
        from .training import warp_hidden_state
        return warp_hidden_state(image_src, depth_dst, None, src_trans_dst, camera_matrix, float("-inf"))
    out = ops.hidden_warp(ops.to_nhwc(image_src, "image_src"), depth_dst, None, src_trans_dst, camera_matrix, float("-inf"))
    return ops.to_api(out)


# ------------------------------------------------------------------------------------------------ glue (cold)
def pose_distance(reference_pose, measurement_pose):
    """Reference utils.py:17-31 (numpy, host)."""
    rel = np.linalg.inv(reference_pose) @ measurement_pose
    R, t = rel[:3, :3], rel[:3, 3]
    R_measure = np.sqrt(2 * (1 - min(3.0, float(np.trace(R))) / 3))
    t_measure = float(np.linalg.norm(t))
    return np.sqrt(t_measure ** 2 + R_measure ** 2), R_measure, t_measure


def is_pose_available(pose):
    """Reference utils.py:261-268."""
    return bool(np.all(np.isfinite(pose)))


class InferenceTimer:
    """Reference utils.py:369-402: CUDA-event pair per keyframe on the current stream, first n_skip samples dropped."""

    def __init__(self, n_skip=20):
        self.times = []
        self.n_skip = n_skip
        self.forward_pass_start = torch.cuda.Event(enable_timing=True)
        self.forward_pass_end = torch.cuda.Event(enable_timing=True)

    def record_start_time(self):
        self.forward_pass_start.record()

    def record_end_time_and_elapsed_time(self):
        self.forward_pass_end.record()
        torch.cuda.synchronize()
        self.times.append(self.forward_pass_start.elapsed_time(self.forward_pass_end))

    def print_statistics(self):
        times = np.array(self.times[self.n_skip:])
        if len(times) == 0:
            print("Not enough time measurements are taken!")
            return
        print("Number of Forward Passes:", len(times))
        for label, fn in (("Mean", np.mean), ("Std", np.std), ("Median", np.median), ("Min", np.min), ("Max", np.max)):
            print("--- %s Inference Time:" % label, fn(times))


def save_predictions(predictions, system_name, scene_name, save_folder):
    np.savez_compressed(os.path.join(str(save_folder), system_name + "_predictions_" + scene_name), predictions)


def save_results(predictions, groundtruths, system_name, scene_name, save_folder, max_depth=np.inf):
    """Reference utils.py:330-348: same .npz names / contents (the on-disk format the TSDF script consumes)."""
    os.makedirs(str(save_folder), exist_ok=True)
    if groundtruths is not None:
        errors = np.array([compute_errors(groundtruths[i], p, max_depth) for i, p in enumerate(predictions)])
        names = ['abs_error', 'abs_relative_error', 'abs_inverse_error', 'squared_relative_error', 'rmse', 'ratio_125',
                 'ratio_125_2', 'ratio_125_3']
        print("Metrics of {} for scene {}:".format(system_name, scene_name))
        print(", ".join("{:>25}".format(n) for n in names))
        print(", ".join("{:25.4f}".format(v) for v in np.nanmean(errors, 0)))
        np.savez_compressed(os.path.join(str(save_folder), system_name + "_errors_" + scene_name), errors)
    save_predictions(np.array(predictions), system_name, scene_name, save_folder)


def visualize_predictions(numpy_reference_image, numpy_measurement_image, numpy_predicted_depth, normalization_mean,
                          normalization_std, normalization_scale, depth_multiplier_for_visualization=5000):
    """Reference utils.py:351-366 (cv2.imshow); needs a display."""
    import cv2

    def to_u8(img):
        return ((img * np.array(normalization_std) + np.array(normalization_mean)) * normalization_scale).astype(np.uint8)

    cv2.imshow("Reference Image", cv2.cvtColor(to_u8(numpy_reference_image), cv2.COLOR_RGB2BGR))
    cv2.imshow("A Measurement Image", cv2.cvtColor(to_u8(numpy_measurement_image), cv2.COLOR_RGB2BGR))
    cv2.imshow("Predicted Depth", (depth_multiplier_for_visualization * numpy_predicted_depth).astype(np.uint16))
    cv2.waitKey()
