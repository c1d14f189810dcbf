"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A CPU (torch fp32) restatement of the reference's plane-sweep depth-inference path, written from the
reference's algorithm (not its text), every function citing the reference file:line it follows
(paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and only as the checker / the timed CPU baseline.

Parity status: PINNED.  This restatement is checked (tests/test_oracle_golden.py) against
  * tests/golden/ops_*.npz / modules_*.npz -- outputs of the UNMODIFIED reference modules imported from
    /root/reference in the build container (generator: oracle/make_golden.py), and
  * the reference's own shipped end-to-end golden
    sample-data/predictions/keyframe_hololens-dataset_320_256_3_dvmvs_fusionnet_online_predictions_000.npz
    (first frames, fixture scene 000, shipped fusionnet weights).

All modules are functional: they take a state dict with the reference's key names.
Geometry is restated with explicit gathers (no grid_sample) so the sampling convention is visible.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


# --------------------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------------------
def get_warp_grid_for_cost_volume_calculation(width, height, device="cpu"):
    """dvmvs/utils.py:34-42 -- homogeneous pixel grid (3, h*w): rows x, y, 1 (row-major pixels)."""
    ys, xs = torch.meshgrid(torch.arange(int(height), dtype=torch.float32),
                            torch.arange(int(width), dtype=torch.float32), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(int(height) * int(width))], dim=0).to(device)


def bilinear_sample_zeros(image, xs, ys):
    """Bilinear gather at PIXEL coordinates (xs, ys) (B,h,w) from image (B,C,H,W), zeros outside --
    the arithmetic of grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) once the
    normalised grid has been mapped back to pixels (dvmvs/utils.py:75-79, :258)."""
    B, C, H, W = image.shape
    x0 = torch.floor(xs)
    y0 = torch.floor(ys)
    out = torch.zeros(B, C, *xs.shape[1:], dtype=image.dtype, device=image.device)
    flat = image.reshape(B, C, H * W)
    for dy in (0, 1):
        for dx in (0, 1):
            xi = x0 + dx
            yi = y0 + dy
            wx = (xs - x0) if dx else (x0 + 1 - xs)
            wy = (ys - y0) if dy else (y0 + 1 - ys)
            valid = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
            xi_c = xi.clamp(0, W - 1).long()
            yi_c = yi.clamp(0, H - 1).long()
            idx = (yi_c * W + xi_c).reshape(B, 1, -1).expand(B, C, -1)
            tap = torch.gather(flat, 2, idx).reshape(B, C, *xs.shape[1:])
            w = torch.where(valid, wx * wy, torch.zeros_like(wx))
            w = torch.nan_to_num(w, nan=0.0, posinf=0.0, neginf=0.0)
            out = out + tap * w.unsqueeze(1)
    return out


def calculate_cost_volume_by_warping(image1, image2, pose1, pose2, K, warp_grid, min_depth, max_depth,
                                     n_depth_levels, device="cpu", dot_product=True):
    """dvmvs/utils.py:45-86.  image1/2 (B,C,h,w); pose cam-to-world (B,4,4);
    K half-resolution intrinsics (B,3,3).  Plane i has inverse depth 1/max + i*step (i=0 farthest)."""
    B, C, h, w = image1.shape
    E = torch.inverse(pose2).bmm(pose1)                       # :51
    R = E[:, 0:3, 0:3]
    t = E[:, 0:3, 3].unsqueeze(-1)
    Kt = K.bmm(t)                                             # :55
    G = K.bmm(R).bmm(torch.inverse(K))                        # :56
    grid = warp_grid.to(image1.device).unsqueeze(0).expand(B, -1, -1)
    base = G.bmm(grid)                                        # :57  (B,3,h*w)
    inv_base = 1.0 / max_depth                                # :59
    inv_step = (1.0 / min_depth - 1.0 / max_depth) / (n_depth_levels - 1)   # :60
    cost = torch.empty(B, n_depth_levels, h, w, dtype=torch.float32, device=image1.device)
    for i in range(n_depth_levels):
        this_depth = 1 / (inv_base + i * inv_step)            # :66
        q = base + Kt / this_depth                            # :68
        x = q[:, 0] / (q[:, 2] + 1e-8)                        # :70 (no behind-camera test)
        y = q[:, 1] / (q[:, 2] + 1e-8)
        gx = (x - w / 2.0) / (w / 2.0)                        # :72-73
        gy = (y - h / 2.0) / (h / 2.0)
        # align_corners=True un-normalisation: pixel = (g + 1)/2 * (size - 1)  => the (w-1)/w shrink quirk
        xs = ((gx + 1.0) / 2.0 * (w - 1)).reshape(B, h, w)
        ys = ((gy + 1.0) / 2.0 * (h - 1)).reshape(B, h, w)
        warped = bilinear_sample_zeros(image2, xs, ys)
        if dot_product:
            cost[:, i] = torch.sum(image1 * warped, dim=1) / C          # :82
        else:
            cost[:, i] = torch.sum(torch.abs(image1 - warped), dim=1)   # :84
    return cost


def cost_volume_fusion(image1, image2s, pose1, pose2s, K, warp_grid, min_depth, max_depth, n_depth_levels,
                       device="cpu", dot_product=True):
    """dvmvs/utils.py:89-107 -- accumulate over measurement frames in list order, divide by M."""
    fused = torch.zeros(image1.shape[0], n_depth_levels, image1.shape[2], image1.shape[3], dtype=torch.float32, device=image1.device)
    for pose2, image2 in zip(pose2s, image2s):
        fused += calculate_cost_volume_by_warping(image1, image2, pose1, pose2, K, warp_grid, min_depth,
                                                  max_depth, n_depth_levels, device, dot_product)
    fused /= len(pose2s)
    return fused


def _unproject(depth, K):
    """kornia 0.3.2 depth_to_3d(normalize_points=False): (B,1,H,W) -> (B,H,W,3)."""
    B, _, H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=depth.device), torch.arange(W, dtype=torch.float32, device=depth.device),
                            indexing="ij")
    fx, fy = K[:, 0, 0].view(B, 1, 1), K[:, 1, 1].view(B, 1, 1)
    cx, cy = K[:, 0, 2].view(B, 1, 1), K[:, 1, 2].view(B, 1, 1)
    d = depth[:, 0]
    return torch.stack([(xs - cx) / fx * d, (ys - cy) / fy * d, d], dim=-1)


def _transform(T, pts):
    """kornia transform_points for a rigid 4x4 (w == 1 so the homogeneous divide is the identity)."""
    R = T[:, None, None, 0:3, 0:3]
    t = T[:, None, None, 0:3, 3]
    return (R @ pts.unsqueeze(-1)).squeeze(-1) + t


def _project(pts, K):
    """kornia project_points: divide by z only where |z| > 1e-8 (else leave x, y), then apply K."""
    z = pts[..., 2]
    scale = torch.where(z.abs() > 1e-8, 1.0 / z, torch.ones_like(z))
    B = pts.shape[0]
    shp = (B,) + (1,) * (pts.dim() - 2)
    u = pts[..., 0] * scale * K[:, 0, 0].view(shp) + K[:, 0, 2].view(shp)
    v = pts[..., 1] * scale * K[:, 1, 1].view(shp) + K[:, 1, 2].view(shp)
    return u, v


def warp_frame_depth(image_src, depth_dst, src_trans_dst, camera_matrix, normalize_points=False,
                     sampling_mode="bilinear"):
    """dvmvs/utils.py:205-258.  normalize_pixel_coordinates + align_corners=True
    samples exactly at the projected pixel (no shrink)."""
    pts = _unproject(depth_dst, camera_matrix)                 # :241
    pts = _transform(src_trans_dst, pts)                       # :247
    pts = torch.stack([pts[..., 0], pts[..., 1], torch.relu(pts[..., 2])], dim=-1)   # :248
    u, v = _project(pts, camera_matrix)                        # :252
    return bilinear_sample_zeros(image_src, u, v)              # :256-258


def get_non_differentiable_rectangle_depth_estimation(reference_pose_torch, measurement_pose_torch,
                                                      previous_depth_torch, full_K_torch, half_K_torch,
                                                      original_width, original_height):
    """dvmvs/utils.py:110-154.  Sorting z descending and keeping the first hit per
    half-res pixel == keeping the maximum (relu'd) z per pixel; unfilled pixels stay 0."""
    B = reference_pose_torch.shape[0]
    hw, hh = int(original_width / 2), int(original_height / 2)
    T = torch.bmm(torch.inverse(reference_pose_torch), measurement_pose_torch)      # :121
    pts = _transform(T, _unproject(previous_depth_torch, full_K_torch)).reshape(B, -1, 3)   # :122-126
    z = torch.relu(pts[:, :, 2])                                                    # :129
    u, v = _project(pts, half_K_torch)          # projection uses the UN-relu'd z (:134-136 gathers pts, not z)
    pu = torch.round(u).long()                                                      # :136 half-to-even
    pv = torch.round(v).long()
    valid = (pu >= 0) & (pv >= 0) & (pu < hw) & (pv < hh)                           # :137-139
    out = np.zeros((B, 1, hh, hw), dtype=np.float32)
    for b in range(B):
        idx = (pv[b][valid[b]] * hw + pu[b][valid[b]]).cpu().numpy()       # the reference syncs to the host here too (utils.py:148)
        zs = z[b][valid[b]].cpu().numpy()
        flat = np.full(hh * hw, -1.0, dtype=np.float32)
        np.maximum.at(flat, idx, zs)
        flat[flat < 0] = 0.0
        out[b, 0] = flat.reshape(hh, hw)
    return torch.from_numpy(out).to(previous_depth_torch.device)


# --------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------
BN_TRAINING = False      # True inside train_mode(): BatchNorm normalises with the batch statistics (nn.Module.train())


class train_mode:
    """Context manager: the functional modules below behave like the reference's nn.Modules after .train()
    (dvmvs/train.py:10-15: batch statistics in every BatchNorm; running statistics are not tracked here)."""

    def __enter__(self):
        global BN_TRAINING
        self._saved, BN_TRAINING = BN_TRAINING, True

    def __exit__(self, *exc):
        global BN_TRAINING
        BN_TRAINING = self._saved


def _bn(sd, p, x):
    if BN_TRAINING:
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, BN_EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _conv_bn_relu(sd, p, x, stride=1, relu=True):
    """dvmvs/layers.py:39-59 conv_layer: Conv(bias=False, pad=(k-1)//2) [+ BN(eps 1e-5) + ReLU].
    `p` is the Sequential prefix: p.0 = conv, p.1 = BN."""
    w = sd[p + ".0.weight"]
    x = F.conv2d(x, w, None, stride, (w.shape[-1] - 1) // 2)
    if (p + ".1.weight") in sd:
        x = _bn(sd, p + ".1", x)
        if relu:
            x = F.relu(x)
    return x


def _depth_head(sd, p, x):
    """dvmvs/layers.py:62-65 depth_layer_3x3: Conv(bias, 3x3, pad 1) + Sigmoid."""
    return torch.sigmoid(F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], 1, 1))


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


# MnasNet-1.0 layers[8:14] (torchvision mnasnet.py): (prefix, stride of first block, n blocks)
_MNAS_STACKS = {
    "layer2": [("layer2.0", 2, 3)],
    "layer3": [("layer3.0", 2, 3)],
    "layer4": [("layer4.0", 2, 3), ("layer4.1", 1, 2)],
    "layer5": [("layer5.0", 2, 4), ("layer5.1", 1, 1)],
}


def _inverted_residual(sd, p, x, stride):
    """torchvision _InvertedResidual: pw-expand+BN+ReLU -> dw(k, pad k//2, stride)+BN+ReLU -> pw-linear+BN,
    residual add iff in == out and stride == 1."""
    w0, w3, w6 = sd[p + ".layers.0.weight"], sd[p + ".layers.3.weight"], sd[p + ".layers.6.weight"]
    y = F.relu(_bn(sd, p + ".layers.1", F.conv2d(x, w0)))
    y = F.relu(_bn(sd, p + ".layers.4", F.conv2d(y, w3, None, stride, w3.shape[-1] // 2, 1, w3.shape[0])))
    y = _bn(sd, p + ".layers.7", F.conv2d(y, w6))
    if stride == 1 and w0.shape[1] == w6.shape[0]:
        y = y + x
    return y


def feature_extractor(sd, image):
    """dvmvs/fusionnet/model.py:122-140 = torchvision mnasnet1_0().layers[0:14] split into 5 outputs."""
    x = F.relu(_bn(sd, "layer1.1", F.conv2d(image, sd["layer1.0.weight"], None, 2, 1)))
    x = F.relu(_bn(sd, "layer1.4", F.conv2d(x, sd["layer1.3.weight"], None, 1, 1, 1, x.shape[1])))
    x = _bn(sd, "layer1.7", F.conv2d(x, sd["layer1.6.weight"]))
    outs = [x]
    for name in ("layer2", "layer3", "layer4", "layer5"):
        for prefix, stride, n in _MNAS_STACKS[name]:
            for i in range(n):
                x = _inverted_residual(sd, "%s.%d" % (prefix, i), x, stride if i == 0 else 1)
        outs.append(x)
    return tuple(outs)


def feature_shrinker(sd, layer1, layer2, layer3, layer4, layer5):
    """dvmvs/fusionnet/model.py:143-164 = torchvision FeaturePyramidNetwork([16,24,40,96,320] -> 32), top-down
    with nearest upsampling; returns levels 1-4 (level-5 output is computed by the reference and dropped)."""
    feats = [layer1, layer2, layer3, layer4, layer5]

    def inner(i, x):
        return F.conv2d(x, sd["fpn.inner_blocks.%d.weight" % i], sd["fpn.inner_blocks.%d.bias" % i])

    def layer(i, x):
        return F.conv2d(x, sd["fpn.layer_blocks.%d.weight" % i], sd["fpn.layer_blocks.%d.bias" % i], 1, 1)

    last = inner(4, feats[4])
    outs = [None] * 5
    outs[4] = layer(4, last)
    for i in (3, 2, 1, 0):
        lat = inner(i, feats[i])
        last = lat + F.interpolate(last, size=lat.shape[-2:], mode="nearest")
        outs[i] = layer(i, last)
    return outs[0], outs[1], outs[2], outs[3]


def cost_volume_encoder(sd, features_half, features_quarter, features_one_eight, features_one_sixteen, cost_volume):
    """dvmvs/fusionnet/model.py:167-224."""
    def block(p, x):
        x = _conv_bn_relu(sd, p + ".down_convolution.down_conv", x, stride=2)
        x = _conv_bn_relu(sd, p + ".standard_convolution.conv1", x)
        return _conv_bn_relu(sd, p + ".standard_convolution.conv2", x)

    inp0 = _conv_bn_relu(sd, "aggregator0", torch.cat([features_half, cost_volume], dim=1))
    out0 = block("encoder_block0", inp0)
    inp1 = _conv_bn_relu(sd, "aggregator1", torch.cat([features_quarter, out0], dim=1))
    out1 = block("encoder_block1", inp1)
    inp2 = _conv_bn_relu(sd, "aggregator2", torch.cat([features_one_eight, out1], dim=1))
    out2 = block("encoder_block2", inp2)
    inp3 = _conv_bn_relu(sd, "aggregator3", torch.cat([features_one_sixteen, out2], dim=1))
    out3 = block("encoder_block3", inp3)
    return inp0, inp1, inp2, inp3, out3


def cost_volume_decoder(sd, image, skip0, skip1, skip2, skip3, bottom, min_depth=0.25, max_depth=20.0):
    """dvmvs/fusionnet/model.py:227-305 (Config.train_min/max_depth = 0.25/20, config.py:7-8)."""
    base = 1 / max_depth
    mult = 1 / min_depth - 1 / max_depth

    def block(p, x, skip, depth):
        x = _conv_bn_relu(sd, p + ".up_convolution.conv", _up2(x))
        x = torch.cat([x, skip] if depth is None else [x, skip, _up2(depth)], dim=1)
        x = _conv_bn_relu(sd, p + ".convolution1", x)
        return _conv_bn_relu(sd, p + ".convolution2", x)

    d1 = block("decoder_block1", bottom, skip3, None)
    s16 = _depth_head(sd, "depth_layer_one_sixteen", d1)
    d2 = block("decoder_block2", d1, skip2, s16)
    s8 = _depth_head(sd, "depth_layer_one_eight", d2)
    d3 = block("decoder_block3", d2, skip1, s8)
    s4 = _depth_head(sd, "depth_layer_quarter", d3)
    d4 = block("decoder_block4", d3, skip0, s4)
    s2 = _depth_head(sd, "depth_layer_half", d4)
    x = torch.cat([_up2(d4), _up2(s2), image], dim=1)
    x = _conv_bn_relu(sd, "refine.1", _conv_bn_relu(sd, "refine.0", x))
    s1 = _depth_head(sd, "depth_layer_full", x)
    return tuple(1.0 / (mult * s + base).squeeze(1) for s in (s1, s2, s4, s8, s16))


def lstm_fusion(sd, current_encoding, current_state, previous_pose, current_pose, estimated_current_depth,
                camera_matrix):
    """dvmvs/fusionnet/model.py:321-337 + dvmvs/convlstm.py:26-64.  Gate order i,f,o,g; LN over (h,w) per
    channel (biased variance, eps 1e-5, no affine); CELU alpha=1."""
    B, C, h, w = current_encoding.shape
    if current_state is None:
        h_cur = torch.zeros(B, C, h, w, device=current_encoding.device)
        c_cur = torch.zeros(B, C, h, w, device=current_encoding.device)
    else:
        h_cur, c_cur = current_state
    if previous_pose is not None:
        T = torch.bmm(torch.inverse(previous_pose), current_pose)                     # convlstm.py:30
        non_valid = estimated_current_depth <= 0.01                                   # :32
        h_cur = warp_frame_depth(h_cur, estimated_current_depth, T, camera_matrix)    # :33-38
        # :39-41 `h_cur.data[non_valid] = 0.0` -- a .data write autograd does not see: the VALUE is zeroed, the
        # gradient passes through the masked positions into the warp as if they were not masked.  Same here.
        h_cur = h_cur - (h_cur * non_valid.expand_as(h_cur).to(h_cur.dtype)).detach()
    cc = F.conv2d(torch.cat([current_encoding, h_cur], dim=1), sd["lstm_cell.conv.weight"], None, 1, 1)   # :43-44
    return lstm_gate_epilogue(cc, c_cur)


def lstm_gate_epilogue(combined_conv, c_cur):
    """dvmvs/convlstm.py:45-59: the gate arithmetic after the convolution (differentiable: the training tests take
    torch autograd through this restatement as the reference gradient)."""
    B, C4, h, w = combined_conv.shape
    C = C4 // 4
    cc_i, cc_f, cc_o, cc_g = torch.split(combined_conv, C, dim=1)                     # :45
    i, f, o = torch.sigmoid(cc_i), torch.sigmoid(cc_f), torch.sigmoid(cc_o)
    g = torch.celu(torch.layer_norm(cc_g, [h, w]))                                    # :52-53
    c_next = torch.layer_norm(f * c_cur + i * g, [h, w])                              # :55-56
    h_next = o * torch.celu(c_next)                                                   # :57
    return h_next, c_next


# --------------------------------------------------------------------------------------------------
# training loss (dvmvs/losses.py), restated; differentiable
# --------------------------------------------------------------------------------------------------
def calculate_loss(groundtruth, prediction):
    """dvmvs/losses.py:53-82: sums over the pixels whose nearest-down-sampled ground truth is non-zero.
    Returns (l1, huber, l1_inv, l1_rel, valid_count)."""
    B, H, W = groundtruth.shape
    _, hs, ws = prediction.shape
    gt = F.interpolate(groundtruth.view(B, 1, H, W), size=(hs, ws), mode="nearest")   # :61-63
    pred = prediction.view(B, 1, hs, ws)
    valid = gt != 0                                                                    # :65
    count = int(valid.sum())
    g, p = gt[valid], pred[valid]
    diff = torch.abs(g - p)                                                            # :73
    huber = F.smooth_l1_loss(p, g, reduction="none").sum()                             # :75-76
    return diff.sum(), huber, torch.abs(1.0 / g - 1.0 / p).sum(), (diff / g).sum(), count   # :78-82


def update_losses(predictions, weights, groundtruth, loss_type):
    """dvmvs/losses.py:26-40, is_training branch: sum_j weights[j] * loss_j / valid_count_j.
    Returns (optimizer_loss, per-scale (n,5) sums as python floats)."""
    column = {"L1": 0, "Huber": 1, "L1-inv": 2, "L1-rel": 3}[loss_type]
    total, sums = 0, []
    for wgt, prediction in zip(weights, predictions):
        parts = calculate_loss(groundtruth, prediction)
        total = total + wgt * (parts[column] / parts[4])
        sums.append([float(v.detach()) if torch.is_tensor(v) else float(v) for v in parts])
    return total, sums


def fusionnet_training_forward(weights, images, depths, poses, K, min_depth=0.25, max_depth=20.0, n_depth_levels=64,
                               loss_type="L1-inv"):
    """dvmvs/fusionnet/run-training.py:183-281 forward_pass (is_training=True), restated: features of every frame of the
    subsequence, then for i >= 1 the pair (reference i, measurement i-1): single-frame cost volume, encoder, ConvLSTM with
    the hidden state warped by the GROUND-TRUTH depth of the reference frame (nearest 1/32, :245-249), decoder, five-scale
    loss with unit weights (:270-278).  images / depths / poses: lists over the subsequence of (B,3,H,W) / (B,H,W) /
    (B,4,4).  Returns the optimizer loss (differentiable w.r.t. `weights`).  Call under train_mode()."""
    B, _, H, W = images[0].shape
    half_K = K.clone()
    half_K[:, 0:2, :] = half_K[:, 0:2, :] * 0.5                                        # :191-192 (scaling = 0.5)
    lstm_K = K.clone()
    lstm_K[:, 0:2, :] = lstm_K[:, 0:2, :] / 32.0                                       # :194-195
    grid = get_warp_grid_for_cost_volume_calculation(W // 2, H // 2)
    feats = [feature_shrinker(weights["fpn"], *feature_extractor(weights["fe"], im)) for im in images]   # :206-215
    loss, state = 0, None
    for i in range(1, len(images)):
        f2, f4, f8, f16 = feats[i]
        cv = calculate_cost_volume_by_warping(f2, feats[i - 1][0], poses[i], poses[i - 1], half_K, grid, min_depth, max_depth,
                                              n_depth_levels, "cpu", True)             # :231-241
        s0, s1, s2, s3, bottom = cost_volume_encoder(weights["cve"], f2, f4, f8, f16, cv)
        de = F.interpolate(depths[i].view(B, 1, H, W), scale_factor=1.0 / 32.0, mode="nearest")          # :249-253
        state = lstm_fusion(weights["lstm"], bottom, state, poses[i - 1], poses[i], de, lstm_K)           # :255-260
        full, half, quarter, eighth, sixteenth = cost_volume_decoder(weights["cvd"], images[i], s0, s1, s2, s3, state[0],
                                                                      min_depth, max_depth)
        step_loss, _ = update_losses([sixteenth, eighth, quarter, half, full], [1, 1, 1, 1, 1], depths[i], loss_type)
        loss = loss + step_loss
    return loss


# --------------------------------------------------------------------------------------------------
# one keyframe of dvmvs/fusionnet/run-testing.py:145-202 (the caller's loop body, restated)
# --------------------------------------------------------------------------------------------------
class FusionnetState:
    def __init__(self):
        self.lstm_state = None
        self.previous_depth = None
        self.previous_pose = None


def fusionnet_step(weights, state, reference_image, reference_pose, measurement_images, measurement_poses, full_K,
                   min_depth=0.25, max_depth=20.0, n_depth_levels=64):
    """weights: dict name -> state dict for 'fe','fpn','cve','lstm','cvd'.  Tensors are (1,...) batched.
    Returns (depth_full (B,H,W), state)."""
    B, _, H, W = reference_image.shape
    half_K = full_K.clone()
    half_K[:, 0:2, :] = half_K[:, 0:2, :] / 2.0                                       # run-testing.py:145-146
    lstm_K = full_K.clone()
    lstm_K[:, 0:2, :] = lstm_K[:, 0:2, :] / 32.0                                      # :148-149
    meas_half = [feature_shrinker(weights["fpn"], *feature_extractor(weights["fe"], im))[0] for im in measurement_images]
    f2, f4, f8, f16 = feature_shrinker(weights["fpn"], *feature_extractor(weights["fe"], reference_image))
    grid = get_warp_grid_for_cost_volume_calculation(W // 2, H // 2)
    cv = cost_volume_fusion(f2, meas_half, reference_pose, measurement_poses, half_K, grid, min_depth, max_depth,
                            n_depth_levels, "cpu", True)                              # :161-171
    s0, s1, s2, s3, bottom = cost_volume_encoder(weights["cve"], f2, f4, f8, f16, cv)
    if state.previous_depth is not None:
        de = get_non_differentiable_rectangle_depth_estimation(reference_pose, state.previous_pose, state.previous_depth,
                                                               full_K, half_K, W, H)  # :179-186
        de = F.interpolate(de, scale_factor=1.0 / 16.0, mode="nearest")               # :187-189
    else:
        de = torch.zeros(B, 1, H // 32, W // 32, device=reference_image.device)       # :191
    state.lstm_state = lstm_fusion(weights["lstm"], bottom, state.lstm_state, state.previous_pose, reference_pose, de, lstm_K)
    pred = cost_volume_decoder(weights["cvd"], reference_image, s0, s1, s2, s3, state.lstm_state[0], min_depth, max_depth)[0]
    state.previous_depth = pred.view(B, 1, H, W)                                      # :201
    state.previous_pose = reference_pose
    return pred, state


def pairnet_step(weights, reference_image, reference_pose, measurement_images, measurement_poses, full_K,
                 min_depth=0.25, max_depth=20.0, n_depth_levels=64):
    """dvmvs/pairnet/run-testing.py:139-164 -- fusionnet without the recurrent cell."""
    B, _, H, W = reference_image.shape
    half_K = full_K.clone()
    half_K[:, 0:2, :] = half_K[:, 0:2, :] / 2.0
    meas_half = [feature_shrinker(weights["fpn"], *feature_extractor(weights["fe"], im))[0] for im in measurement_images]
    f2, f4, f8, f16 = feature_shrinker(weights["fpn"], *feature_extractor(weights["fe"], reference_image))
    grid = get_warp_grid_for_cost_volume_calculation(W // 2, H // 2)
    cv = cost_volume_fusion(f2, meas_half, reference_pose, measurement_poses, half_K, grid, min_depth, max_depth,
                            n_depth_levels, "cpu", True)
    s0, s1, s2, s3, bottom = cost_volume_encoder(weights["cve"], f2, f4, f8, f16, cv)
    return cost_volume_decoder(weights["cvd"], reference_image, s0, s1, s2, s3, bottom, min_depth, max_depth)[0]


def rel_l1_inverse_depth(pred, gold):
    """The parity metric of BASELINE.json: sum|1/p - 1/g| / sum|1/g|."""
    pred = torch.as_tensor(pred, dtype=torch.float64)
    gold = torch.as_tensor(gold, dtype=torch.float64)
    return float((1.0 / pred - 1.0 / gold).abs().sum() / (1.0 / gold).abs().sum())


# --------------------------------------------------------------------------------------------------
# state-dict shape tables (so tests / bench can synthesise weights without constructing modules)
# --------------------------------------------------------------------------------------------------
def _bn_shapes(d, p, c):
    d[p + ".weight"] = (c,)
    d[p + ".bias"] = (c,)
    d[p + ".running_mean"] = (c,)
    d[p + ".running_var"] = (c,)
    d[p + ".num_batches_tracked"] = ()


def state_dict_shapes(n_depth_levels=64, with_lstm=True):
    """Key -> shape for the five modules; aggregator0 has n_depth_levels + 32 inputs
    (fusionnet/model.py:170)."""
    fe = {}
    fe["layer1.0.weight"] = (32, 3, 3, 3)
    _bn_shapes(fe, "layer1.1", 32)
    fe["layer1.3.weight"] = (32, 1, 3, 3)
    _bn_shapes(fe, "layer1.4", 32)
    fe["layer1.6.weight"] = (16, 32, 1, 1)
    _bn_shapes(fe, "layer1.7", 16)
    cfg = [("layer2.0", 16, 24, 3, 3, 3), ("layer3.0", 24, 40, 5, 3, 3), ("layer4.0", 40, 80, 5, 6, 3),
           ("layer4.1", 80, 96, 3, 6, 2), ("layer5.0", 96, 192, 5, 6, 4), ("layer5.1", 192, 320, 3, 6, 1)]
    for prefix, cin, cout, k, exp, n in cfg:
        for i in range(n):
            ci = cin if i == 0 else cout
            mid = ci * exp
            p = "%s.%d.layers" % (prefix, i)
            fe[p + ".0.weight"] = (mid, ci, 1, 1)
            _bn_shapes(fe, p + ".1", mid)
            fe[p + ".3.weight"] = (mid, 1, k, k)
            _bn_shapes(fe, p + ".4", mid)
            fe[p + ".6.weight"] = (cout, mid, 1, 1)
            _bn_shapes(fe, p + ".7", cout)
    fpn = {}
    for i, c in enumerate([16, 24, 40, 96, 320]):
        fpn["fpn.inner_blocks.%d.weight" % i] = (32, c, 1, 1)
        fpn["fpn.inner_blocks.%d.bias" % i] = (32,)
    for i in range(5):
        fpn["fpn.layer_blocks.%d.weight" % i] = (32, 32, 3, 3)
        fpn["fpn.layer_blocks.%d.bias" % i] = (32,)

    def cbr(d, p, cin, cout, k):
        d[p + ".0.weight"] = (cout, cin, k, k)
        _bn_shapes(d, p + ".1", cout)

    cve = {}
    hyper = 32
    chans = [hyper, hyper * 2, hyper * 4, hyper * 8, hyper * 16]
    ks = [5, 3, 3, 3]
    for lvl in range(4):
        cin = (n_depth_levels + 32) if lvl == 0 else (chans[lvl] + 32)
        cbr(cve, "aggregator%d" % lvl, cin, chans[lvl], ks[lvl])
        cbr(cve, "encoder_block%d.down_convolution.down_conv" % lvl, chans[lvl], chans[lvl + 1], ks[lvl])
        cbr(cve, "encoder_block%d.standard_convolution.conv1" % lvl, chans[lvl + 1], chans[lvl + 1], ks[lvl])
        cbr(cve, "encoder_block%d.standard_convolution.conv2" % lvl, chans[lvl + 1], chans[lvl + 1], ks[lvl])
    cvd = {}
    for n, (cin, cout, k, plus) in enumerate([(512, 256, 3, 0), (256, 128, 3, 1), (128, 64, 3, 1), (64, 32, 5, 1)], start=1):
        cbr(cvd, "decoder_block%d.up_convolution.conv" % n, cin, cout, k)
        cbr(cvd, "decoder_block%d.convolution1" % n, cin + plus, cout, k)
        cbr(cvd, "decoder_block%d.convolution2" % n, cout, cout, k)
    cbr(cvd, "refine.0", 36, 32, 5)
    cbr(cvd, "refine.1", 32, 32, 5)
    for name, c in [("one_sixteen", 256), ("one_eight", 128), ("quarter", 64), ("half", 32), ("full", 32)]:
        cvd["depth_layer_%s.0.weight" % name] = (1, c, 3, 3)
        cvd["depth_layer_%s.0.bias" % name] = (1,)
    out = {"fe": fe, "fpn": fpn, "cve": cve, "cvd": cvd}
    if with_lstm:
        out["lstm"] = {"lstm_cell.conv.weight": (2048, 1024, 3, 3)}
    return out
