"""Differentiable forms of the geometric / recurrent ops of the training step, each with a
hand-written sm_100a backward kernel behind `torch.autograd.Function`:

    plane_sweep_cost_volume      dvmvs/utils.py:45-107 under autograd (fusionnet/run-training.py:231 calls it per pair)
    warp_hidden_state            dvmvs/utils.py:205-258 + the mask of dvmvs/convlstm.py:32,40-41 (BPTT through the warp)
    lstm_gate_epilogue           dvmvs/convlstm.py:45-59 (gate non-linearities, two LayerNorms, state update)
    multi_scale_depth_loss       dvmvs/losses.py:26-82 update_losses / calculate_loss

`dvmvs.utils.cost_volume_fusion`, `calculate_cost_volume_by_warping` and `warp_frame_depth` route here on their own when
an input requires grad, so a training script keeps calling the reference's names.  What is NOT here: derivatives of the
convolution stack -- the drop-in modules are inference-only (they raise in train() mode); a training step today pairs
these ops with the caller's own convolution layers.  Poses, intrinsics and the (ground-truth) depth used by the hidden
warp get no gradient, as in the reference's training (fixed poses, run-training.py:245-258)."""
import ctypes

import torch

from . import _native as N
from . import _ops as ops

LOSS_TYPES = {"L1": N.LOSS_L1, "L1-inv": N.LOSS_L1_INV, "L1-rel": N.LOSS_L1_REL, "Huber": N.LOSS_HUBER}
_LOSS_COLUMN = {N.LOSS_L1: 0, N.LOSS_HUBER: 1, N.LOSS_L1_INV: 2, N.LOSS_L1_REL: 3}      # column of `sums` each type reads


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


# ---------------------------------------------------------------------------------------------------- plane sweep
class _PlaneSweep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose1, K, min_depth, max_depth, n_depth_levels, M, *rest):
        image1, image2s, pose2s = rest[0], list(rest[1:1 + M]), list(rest[1 + M:1 + 2 * M])
        ref = ops.to_nhwc(image1, "image1")
        meas = [ops.to_nhwc(t, "image2") for t in image2s]
        if ref.shape[-1] != 32:
            raise RuntimeError("differentiable plane sweep: C must be 32 (half-resolution FPN features), got %d" % ref.shape[-1])
        cost = ops.plane_sweep(ref, meas, pose1, pose2s, K, min_depth, max_depth, n_depth_levels, True)
        ctx.save_for_backward(ref, pose1.contiguous(), K.contiguous(), *meas, *[p.contiguous() for p in pose2s])
        ctx.cfg = (float(min_depth), float(max_depth), int(n_depth_levels), int(M))
        ctx.need = (ctx.needs_input_grad[6], [ctx.needs_input_grad[7 + m] for m in range(M)])
        return ops.to_api(cost)

    @staticmethod
    def backward(ctx, grad_cost):
        min_depth, max_depth, D, M = ctx.cfg
        saved = ctx.saved_tensors
        ref, pose1, K, meas, pose2s = saved[0], saved[1], saved[2], list(saved[3:3 + M]), list(saved[3 + M:3 + 2 * M])
        B, h, w, C = ref.shape
        g = ops.to_nhwc(grad_cost, "grad_cost")
        g_ref = torch.empty_like(ref)
        g_meas = [torch.empty_like(t) for t in meas]
        N.check(N.lib().dvmvs_plane_sweep_backward(ref.data_ptr(), _ptr_array(meas), pose1.data_ptr(), _ptr_array(pose2s), K.data_ptr(),
                                                   g.data_ptr(), g_ref.data_ptr(), _ptr_array(g_meas), B, C, h, w, D, M, min_depth, max_depth,
                                                   N.SWEEP_DOT, ops._stream()), "plane_sweep_backward")
        need_ref, need_meas = ctx.need
        grads = [ops.to_api(g_ref) if need_ref else None] + [ops.to_api(t) if n else None for t, n in zip(g_meas, need_meas)]
        return (None, None, None, None, None, None, *grads, *([None] * M))


def plane_sweep_cost_volume(image1, image2s, pose1, pose2s, K, min_depth, max_depth, n_depth_levels):
    """Differentiable cost_volume_fusion (dot-product cost): (B,32,h,w) features -> (B,D,h,w), gradients to image1 and
    every image2s[m]."""
    image2s, pose2s = list(image2s), list(pose2s)
    if len(image2s) != len(pose2s) or not image2s:
        raise ValueError("need as many measurement poses as measurement images (>= 1)")
    return _PlaneSweep.apply(pose1, K, min_depth, max_depth, n_depth_levels, len(image2s), image1, *image2s, *pose2s)


# ---------------------------------------------------------------------------------------------------- hidden-state warp
class _HiddenWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_src, depth, prev_pose, cur_pose, K, invalid_thresh):
        h = ops.to_nhwc(image_src, "image_src")
        out = ops.hidden_warp(h, depth, prev_pose, cur_pose, K, invalid_thresh)
        ctx.has_prev = prev_pose is not None
        keep = [depth.contiguous(), cur_pose.contiguous(), K.contiguous()] + ([prev_pose.contiguous()] if ctx.has_prev else [])
        ctx.save_for_backward(*keep)
        ctx.thresh = float(invalid_thresh)
        return ops.to_api(out)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        depth, cur_pose, K = saved[0], saved[1], saved[2]
        prev_ptr = saved[3].data_ptr() if ctx.has_prev else None
        g = ops.to_nhwc(grad_out, "grad_out")
        B, h, w, C = g.shape
        g_in = torch.empty_like(g)
        # The reference masks with `h_cur.data[non_valid] = 0.0` (convlstm.py:41): a .data write is invisible to autograd,
        # so its gradient flows through the masked positions into the warp as if they were unmasked.  Reproduced: the
        # backward kernel runs without the mask (ctx.thresh only shaped the forward value).
        N.check(N.lib().dvmvs_hidden_warp_backward(g.data_ptr(), depth.data_ptr(), prev_ptr, cur_pose.data_ptr(), K.data_ptr(), g_in.data_ptr(),
                                                   B, C, h, w, float("-inf"), ops._stream()), "hidden_warp_backward")
        return ops.to_api(g_in), None, None, None, None, None


def warp_hidden_state(image_src, depth_dst, previous_pose, current_pose, camera_matrix, invalid_thresh=0.01):
    """convlstm.py:30-41: warp `image_src` (B,C,h,w) with transformation inverse(previous_pose) @ current_pose and zero
    the positions whose depth is <= invalid_thresh.  previous_pose=None: `current_pose` is the ready-made src_trans_dst
    (plain warp_frame_depth; pass invalid_thresh=float('-inf') for no mask).  Gradient to image_src only; as in the
    reference the mask does not act on the gradient (see _HiddenWarp.backward)."""
    return _HiddenWarp.apply(image_src, depth_dst, previous_pose, current_pose, camera_matrix, invalid_thresh)


# ---------------------------------------------------------------------------------------------------- ConvLSTM gates
class _LstmGates(torch.autograd.Function):
    @staticmethod
    def forward(ctx, combined_conv, c_cur):
        gates = ops.to_nhwc(combined_conv, "combined_conv")
        c = ops.to_nhwc(c_cur, "c_cur")
        if gates.shape[-1] != 4 * c.shape[-1]:
            raise ValueError("combined_conv must have 4x the channels of c_cur (i,f,o,g)")
        h_next, c_next = ops.lstm_gates(gates, c)
        ctx.save_for_backward(gates, c)
        return ops.to_api(h_next), ops.to_api(c_next)

    @staticmethod
    def backward(ctx, grad_h, grad_c):
        gates, c = ctx.saved_tensors
        B, h, w, C = c.shape
        gh = ops.to_nhwc(grad_h, "grad_h") if grad_h is not None else torch.zeros_like(c)
        gc_ptr = None
        if grad_c is not None:
            gc = ops.to_nhwc(grad_c, "grad_c")
            gc_ptr = gc.data_ptr()
        g_gates = torch.empty_like(gates)
        g_c = torch.empty_like(c)
        N.check(N.lib().dvmvs_lstm_gates_backward(gates.data_ptr(), c.data_ptr(), gh.data_ptr(), gc_ptr, g_gates.data_ptr(), g_c.data_ptr(),
                                                  B, h, w, C, ops._stream()), "lstm_gates_backward")
        return ops.to_api(g_gates), ops.to_api(g_c)


def lstm_gate_epilogue(combined_conv, c_cur):
    """convlstm.py:45-59 after the gate convolution: combined_conv (B,4*C,h,w) split i,f,o,g; returns (h_next, c_next)."""
    return _LstmGates.apply(combined_conv, c_cur)


# ---------------------------------------------------------------------------------------------------- loss
class _DepthLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, groundtruth, weights, loss_type, *predictions):
        n = len(predictions)
        gt = ops.require_cuda_f32(groundtruth, "groundtruth").contiguous()
        B, H, W = gt.shape
        preds = [ops.require_cuda_f32(p, "prediction").contiguous() for p in predictions]
        for p in preds:
            if p.dim() != 3 or p.shape[0] != B:
                raise ValueError("predictions must be (B, h, w) tensors, got %s" % (tuple(p.shape),))
        hs = (ctypes.c_int * n)(*[p.shape[1] for p in preds])
        ws = (ctypes.c_int * n)(*[p.shape[2] for p in preds])
        sums = (torch.zeros if N.DRYRUN else torch.empty)((n, 5), dtype=torch.float32, device=gt.device)   # zeroed by the entry point
        N.check(N.lib().dvmvs_depth_loss_forward(_ptr_array(preds), hs, ws, n, gt.data_ptr(), sums.data_ptr(), B, H, W, ops._stream()),
                "depth_loss_forward")
        ctx.save_for_backward(gt, sums, *preds)
        ctx.cfg = (n, [float(x) for x in weights], int(loss_type))
        ratio = sums[:, _LOSS_COLUMN[int(loss_type)]] / sums[:, 4]         # per-scale loss / valid count, on the device
        loss = ratio[0] * float(weights[0])
        for j in range(1, n):
            loss = loss + ratio[j] * float(weights[j])
        ctx.mark_non_differentiable(sums)
        return loss, sums

    @staticmethod
    def backward(ctx, grad_loss, _grad_sums):
        n, weights, loss_type = ctx.cfg
        saved = ctx.saved_tensors
        gt, sums, preds = saved[0], saved[1], list(saved[2:])
        B, H, W = gt.shape
        hs = (ctypes.c_int * n)(*[p.shape[1] for p in preds])
        ws = (ctypes.c_int * n)(*[p.shape[2] for p in preds])
        wt = (ctypes.c_float * n)(*weights)
        grads = [torch.empty_like(p) for p in preds]
        up = grad_loss.to(torch.float32).contiguous()
        N.check(N.lib().dvmvs_depth_loss_backward(_ptr_array(preds), _ptr_array(grads), hs, ws, wt, n, gt.data_ptr(), sums.data_ptr(),
                                                  up.data_ptr(), loss_type, B, H, W, ops._stream()), "depth_loss_backward")
        return (None, None, None, *grads)


def multi_scale_depth_loss(predictions, weights, groundtruth, loss_type="L1-inv"):
    """losses.py:26-40 (is_training branch): sum_j weights[j] * loss_j / valid_count_j over the prediction scales, every
    scale against the nearest-down-sampled ground truth, in one launch.  Returns (optimizer_loss, sums) where sums (n,5)
    holds per scale [l1, huber, l1_inv, l1_rel, valid_count] -- what calculate_loss returns and the LossMeters consume
    (losses.py:43-46), still on the device."""
    if loss_type not in LOSS_TYPES:
        raise ValueError("loss_type must be one of %s" % sorted(LOSS_TYPES))
    predictions = list(predictions)
    if len(predictions) != len(weights) or not predictions:
        raise ValueError("need one weight per prediction (>= 1)")
    return _DepthLoss.apply(groundtruth, list(weights), LOSS_TYPES[loss_type], *predictions)
