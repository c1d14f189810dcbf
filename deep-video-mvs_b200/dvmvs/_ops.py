"""Tensor-level wrappers over the C ABI.  Internal activations are channel-last fp32 torch tensors of shape
(B, H, W, C); the reference-facing API hands out / accepts (B, C, H, W) tensors, which are zero-copy permuted
views of the same storage (torch.channels_last strides).  torch is used for device memory and the current
stream only."""
import ctypes

import torch

from . import _native as N


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t)))
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (the B200 path has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))
    return t


def to_nhwc(x, name="input"):
    """(B,C,H,W) logical tensor -> contiguous (B,H,W,C) tensor (zero-copy when already channels_last)."""
    require_cuda_f32(x, name)
    if x.dim() != 4:
        raise ValueError("%s must have shape (B, C, H, W), got %s" % (name, tuple(x.shape)))
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    B, C, H, W = x.shape
    if C == 1:
        return x.contiguous().view(B, H, W, 1)
    xc = x.contiguous()
    y = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
    N.check(N.lib().dvmvs_nchw_to_nhwc(xc.data_ptr(), y.data_ptr(), B, C, H, W, _stream()), "nchw_to_nhwc")
    return y


def to_api(x_nhwc):
    """(B,H,W,C) contiguous -> (B,C,H,W) view (channels_last strides), no copy."""
    return x_nhwc.permute(0, 3, 1, 2)


def to_nchw_contiguous(x_nhwc):
    B, H, W, C = x_nhwc.shape
    y = torch.empty((B, C, H, W), dtype=torch.float32, device=x_nhwc.device)
    N.check(N.lib().dvmvs_nhwc_to_nchw(x_nhwc.data_ptr(), y.data_ptr(), B, C, H, W, _stream()), "nhwc_to_nchw")
    return y


class PackedConv:
    """Dense conv weights in the kernel layout [k][k][Cin][Cout] with eval-mode BatchNorm folded in
    (scale into the weights, shift into the bias; folded in fp64, stored fp32)."""

    def __init__(self, conv_weight, conv_bias=None, bn=None, stride=1, act=N.ACT_NONE):
        w = conv_weight.detach().to(torch.float64)
        cout = w.shape[0]
        b = conv_bias.detach().to(torch.float64) if conv_bias is not None else None
        if bn is not None:
            scale = bn.weight.detach().to(torch.float64) / torch.sqrt(bn.running_var.detach().to(torch.float64) + bn.eps)
            shift = bn.bias.detach().to(torch.float64) - bn.running_mean.detach().to(torch.float64) * scale
            w = w * scale.view(-1, 1, 1, 1)
            b = shift if b is None else b * scale + shift
        self.ksize = int(w.shape[2])
        self.cin = int(w.shape[1])
        self.cout = int(cout)
        self.stride = int(stride)
        self.act = act
        self.weight = w.permute(2, 3, 1, 0).contiguous().to(torch.float32)          # [k][k][Cin][Cout]
        self.bias = b.to(torch.float32).contiguous() if b is not None else None


class PackedDepthwise:
    def __init__(self, conv_weight, bn, stride, act=N.ACT_RELU):
        w = conv_weight.detach().to(torch.float64)                                  # (C,1,k,k)
        scale = bn.weight.detach().to(torch.float64) / torch.sqrt(bn.running_var.detach().to(torch.float64) + bn.eps)
        shift = bn.bias.detach().to(torch.float64) - bn.running_mean.detach().to(torch.float64) * scale
        w = w * scale.view(-1, 1, 1, 1)
        self.ksize = int(w.shape[2])
        self.channels = int(w.shape[0])
        self.stride = int(stride)
        self.act = act
        self.weight = w[:, 0].permute(1, 2, 0).contiguous().to(torch.float32)       # [k][k][C]
        self.bias = shift.to(torch.float32).contiguous()


def conv2d(sources, pc, residual=None, residual_mode=N.RES_NONE, aux=None):
    """sources: list of (nhwc tensor, mode) with mode SRC_DIRECT / SRC_UPSAMPLE2X (tensor at half resolution).
    Returns out (B,Hout,Wout,Cout) [, aux_out] -- aux = (mult, base) emits 1/(mult*act(y)+base) as well."""
    d = N.ConvDesc()
    first, mode0 = sources[0]
    B = first.shape[0]
    Hin = first.shape[1] * (2 if mode0 == N.SRC_UPSAMPLE2X else 1)
    Win = first.shape[2] * (2 if mode0 == N.SRC_UPSAMPLE2X else 1)
    cin = 0
    for i, (t, mode) in enumerate(sources):
        f = 2 if mode == N.SRC_UPSAMPLE2X else 1
        if t.shape[0] != B or t.shape[1] * f != Hin or t.shape[2] * f != Win:
            raise ValueError("conv2d: source %d has shape %s, expected spatial %dx%d (mode %d)" % (i, tuple(t.shape), Hin, Win, mode))
        d.src[i] = t.data_ptr()
        d.src_channels[i] = t.shape[3]
        d.src_mode[i] = mode
        cin += t.shape[3]
    if cin != pc.cin:
        raise ValueError("conv2d: %d input channels given, weights expect %d" % (cin, pc.cin))
    d.n_src = len(sources)
    pad = (pc.ksize - 1) // 2
    Hout = (Hin + 2 * pad - pc.ksize) // pc.stride + 1
    Wout = (Win + 2 * pad - pc.ksize) // pc.stride + 1
    out = torch.empty((B, Hout, Wout, pc.cout), dtype=torch.float32, device=first.device)
    d.weight = pc.weight.data_ptr()
    d.bias = pc.bias.data_ptr() if pc.bias is not None else None
    d.residual_mode = residual_mode
    if residual is not None:
        d.residual = residual.data_ptr()
        d.Hr, d.Wr = residual.shape[1], residual.shape[2]
    d.out = out.data_ptr()
    aux_out = None
    if aux is not None:
        aux_out = torch.empty_like(out)
        d.aux_out = aux_out.data_ptr()
        d.aux_mult, d.aux_base = aux
    d.B, d.Hin, d.Win, d.Cout = B, Hin, Win, pc.cout
    d.ksize, d.stride, d.act = pc.ksize, pc.stride, pc.act
    N.check(N.lib().dvmvs_conv2d(ctypes.byref(d), _stream()), "conv2d")
    return (out, aux_out) if aux is not None else out


def dwconv2d(x, pd):
    B, H, W, C = x.shape
    pad = pd.ksize // 2
    Hout = (H + 2 * pad - pd.ksize) // pd.stride + 1
    Wout = (W + 2 * pad - pd.ksize) // pd.stride + 1
    y = torch.empty((B, Hout, Wout, C), dtype=torch.float32, device=x.device)
    N.check(N.lib().dvmvs_dwconv2d(x.data_ptr(), pd.weight.data_ptr(), pd.bias.data_ptr(), y.data_ptr(), B, H, W, C,
                                   pd.ksize, pd.stride, pd.act, _stream()), "dwconv2d")
    return y


def plane_sweep(ref_nhwc, meas_nhwc_list, pose1, pose2_list, K, min_depth, max_depth, n_depth_levels, dot_product=True,
                force_generic=False):
    B, h, w, C = ref_nhwc.shape
    M = len(meas_nhwc_list)
    if M < 1 or M != len(pose2_list):
        raise ValueError("plane_sweep: need >= 1 measurement frame and as many poses (got %d, %d)" % (M, len(pose2_list)))
    for m in meas_nhwc_list:
        if tuple(m.shape) != (B, h, w, C):
            raise ValueError("plane_sweep: measurement features %s != reference features %s" % (tuple(m.shape), (B, h, w, C)))
    pose1 = require_cuda_f32(pose1, "pose1").contiguous()
    K = require_cuda_f32(K, "K").contiguous()
    poses = [require_cuda_f32(p, "pose2").contiguous() for p in pose2_list]
    if tuple(pose1.shape) != (B, 4, 4) or tuple(K.shape) != (B, 3, 3) or any(tuple(p.shape) != (B, 4, 4) for p in poses):
        raise ValueError("plane_sweep: poses must be (B,4,4) and K (B,3,3)")
    out = torch.empty((B, h, w, int(n_depth_levels)), dtype=torch.float32, device=ref_nhwc.device)
    meas_ptrs = (ctypes.c_void_p * M)(*[m.data_ptr() for m in meas_nhwc_list])
    pose_ptrs = (ctypes.c_void_p * M)(*[p.data_ptr() for p in poses])
    fn = N.lib().dvmvs_plane_sweep_generic if force_generic else N.lib().dvmvs_plane_sweep_fused
    N.check(fn(ref_nhwc.data_ptr(), meas_ptrs, pose1.data_ptr(), pose_ptrs, K.data_ptr(), out.data_ptr(), B, C, h, w,
               int(n_depth_levels), M, float(min_depth), float(max_depth), N.SWEEP_DOT if dot_product else N.SWEEP_SAD, _stream()),
            "plane_sweep_fused")
    return out


def hidden_warp(h_nhwc, depth_b1hw, prev_pose, cur_pose, K, invalid_thresh):
    B, h, w, C = h_nhwc.shape
    depth = require_cuda_f32(depth_b1hw, "depth").contiguous()
    K = require_cuda_f32(K, "camera_matrix").contiguous()
    cur_pose = require_cuda_f32(cur_pose, "pose").contiguous()
    prev_ptr = None
    if prev_pose is not None:
        prev_pose = require_cuda_f32(prev_pose, "previous_pose").contiguous()
        prev_ptr = prev_pose.data_ptr()
    out = torch.empty_like(h_nhwc)
    N.check(N.lib().dvmvs_hidden_warp(h_nhwc.data_ptr(), depth.data_ptr(), prev_ptr, cur_pose.data_ptr(), K.data_ptr(),
                                      out.data_ptr(), B, C, h, w, float(invalid_thresh), _stream()), "hidden_warp")
    return out


def depth_reproject(cur_pose, prev_pose, prev_depth, full_K, half_K, H, W):
    B = cur_pose.shape[0]
    args = [require_cuda_f32(t, n).contiguous() for t, n in ((cur_pose, "reference_pose"), (prev_pose, "measurement_pose"),
                                                             (prev_depth, "previous_depth"), (full_K, "full_K"), (half_K, "half_K"))]
    out = torch.empty((B, 1, H // 2, W // 2), dtype=torch.float32, device=cur_pose.device)
    N.check(N.lib().dvmvs_depth_reproject(args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), args[3].data_ptr(),
                                          args[4].data_ptr(), out.data_ptr(), B, H, W, _stream()), "depth_reproject")
    return out


def lstm_gates(gates_nhwc, c_nhwc):
    B, h, w, C4 = gates_nhwc.shape
    C = C4 // 4
    h_out = torch.empty((B, h, w, C), dtype=torch.float32, device=gates_nhwc.device)
    c_out = torch.empty_like(h_out)
    N.check(N.lib().dvmvs_lstm_gates(gates_nhwc.data_ptr(), c_nhwc.data_ptr(), h_out.data_ptr(), c_out.data_ptr(), B, h, w, C,
                                     _stream()), "lstm_gates")
    return h_out, c_out


def upsample2x(x_nhwc):
    B, H, W, C = x_nhwc.shape
    y = torch.empty((B, 2 * H, 2 * W, C), dtype=torch.float32, device=x_nhwc.device)
    N.check(N.lib().dvmvs_upsample2x(x_nhwc.data_ptr(), y.data_ptr(), B, H, W, C, _stream()), "upsample2x")
    return y
