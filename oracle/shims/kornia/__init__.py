"""TEST INFRASTRUCTURE ONLY -- stand-in for kornia==0.3.2 (reference README.md:68), which is not
installed in this image.  Used solely by oracle/make_golden.py so that the UNMODIFIED reference under
/root/reference can be imported to generate golden vectors.  Restates the published 0.3.2 semantics of
the four functions the reference calls (dvmvs/utils.py:122,124,136,241,247,252,256).
"""
import torch


def depth_to_3d(depth, camera_matrix, normalize_points=False):
    # kornia 0.3.2 geometry/depth.py: unproject the un-normalised pixel meshgrid, scale by depth
    b, _, h, w = depth.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=depth.dtype, device=depth.device),
                            torch.arange(w, dtype=depth.dtype, device=depth.device), indexing="ij")
    fx = camera_matrix[:, 0, 0].view(b, 1, 1)
    fy = camera_matrix[:, 1, 1].view(b, 1, 1)
    cx = camera_matrix[:, 0, 2].view(b, 1, 1)
    cy = camera_matrix[:, 1, 2].view(b, 1, 1)
    x = (xs.unsqueeze(0) - cx) / fx
    y = (ys.unsqueeze(0) - cy) / fy
    pts = torch.stack([x, y, torch.ones_like(x)], dim=1)
    if normalize_points:
        pts = torch.nn.functional.normalize(pts, dim=1, p=2)
    return pts * depth


def convert_points_from_homogeneous(points, eps=1e-8):
    z = points[..., -1:]
    mask = torch.abs(z) > eps
    scale = torch.ones_like(z).masked_scatter_(mask, torch.tensor(1.0, dtype=z.dtype, device=z.device) / z[mask])
    return scale * points[..., :-1]


def convert_points_to_homogeneous(points):
    return torch.nn.functional.pad(points, [0, 1], "constant", 1.0)


def transform_points(trans_01, points_1):
    points_1_h = convert_points_to_homogeneous(points_1)
    points_0_h = torch.matmul(trans_01.unsqueeze(1), points_1_h.unsqueeze(-1)).squeeze(-1)
    return convert_points_from_homogeneous(points_0_h)


def project_points(point_3d, camera_matrix):
    xy = convert_points_from_homogeneous(point_3d)
    fx = camera_matrix[..., 0, 0]
    fy = camera_matrix[..., 1, 1]
    cx = camera_matrix[..., 0, 2]
    cy = camera_matrix[..., 1, 2]
    u = xy[..., 0] * fx + cx
    v = xy[..., 1] * fy + cy
    return torch.stack([u, v], dim=-1)


def normalize_pixel_coordinates(pixel_coordinates, height, width, eps=1e-8):
    hw = torch.stack([torch.tensor(width), torch.tensor(height)]).to(pixel_coordinates.device).to(pixel_coordinates.dtype)
    factor = torch.tensor(2.0, dtype=pixel_coordinates.dtype, device=pixel_coordinates.device) / (hw - 1).clamp(eps)
    return factor * pixel_coordinates - 1


def _na(*a, **k):
    raise NotImplementedError("training-only kornia colour augmentation is outside the oracle's scope")


adjust_brightness = adjust_gamma = adjust_contrast = _na
