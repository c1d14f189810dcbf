"""TSDF fusion of predicted depth maps under the reference's names: `TSDFVolume` / `TSDFFusion` of
sample-data/run-tsdf-reconstruction.py.  `TSDFVolume.integrate` is one launch of
dvmvs_tsdf_integrate (csrc/tsdf.cu) on volumes resident in HBM; its results are bit-identical to the reference's CPU path
(use_gpu=False -- what the reference runs wherever pycuda is missing).  There is no CPU fallback: without the CUDA library or
a GPU the constructor raises.

Same constructor / integrate / get_volume signatures and argument meaning as the reference (:34, :220, :325).  `integrate`
accepts numpy arrays (as the reference's caller passes, :500-560) or CUDA tensors (depth maps straight from the network:
no host round trip).  Mesh extraction (marching cubes, :329-358) is scikit-image's job in the reference too and is delegated to
it when installed."""
import ctypes

import numpy as np
import torch

from . import _native as N
from ._ops import _stream


class TSDFVolume(object):
    """Volumetric TSDF fusion of RGB-D frames (run-tsdf-reconstruction.py:30-358)."""

    def __init__(self, vol_bnds, voxel_size, use_gpu=True, device=None):
        vol_bnds = np.array(vol_bnds, dtype=np.float64)
        if vol_bnds.shape != (3, 2):
            raise AssertionError("[!] `vol_bnds` should be of shape (3, 2).")          # the reference's message (:43)
        if not use_gpu:
            raise RuntimeError("TSDFVolume: this package has no CPU path (use_gpu=False is the reference's numba fallback)")
        if not N.DRYRUN and not torch.cuda.is_available():
            raise RuntimeError("TSDFVolume: needs a CUDA device (no CPU fallback)")
        self._voxel_size = float(voxel_size)
        self._trunc_margin = 5 * self._voxel_size                                     # :46
        self._color_const = 256 * 256
        self._vol_dim = np.ceil((vol_bnds[:, 1] - vol_bnds[:, 0]) / self._voxel_size).astype(int)     # :50
        vol_bnds[:, 1] = vol_bnds[:, 0] + self._vol_dim * self._voxel_size
        self._vol_bnds = vol_bnds
        self._vol_origin = vol_bnds[:, 0].astype(np.float32)                           # :52
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type == "cuda" and self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        shape = tuple(int(d) for d in self._vol_dim)
        self._tsdf_vol = torch.ones(shape, dtype=torch.float32, device=self.device)     # :57-61
        self._weight_vol = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self._color_vol = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self._updated = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._origin_c = (ctypes.c_float * 3)(*[float(v) for v in self._vol_origin])
        self._staging = {}
        self.gpu_mode = True

    # ---- the hot path -------------------------------------------------------------------------------------------------
    def _upload(self, a, name):
        """Host array -> device through a two-deep ring of pinned staging buffers owned by the volume (allocating pinned
        memory per frame would cost more than the kernel); a slot is reused only after its previous copy has completed."""
        key = (name, tuple(a.shape), a.dtype)
        ring = self._staging.get(key)
        if ring is None:
            ring = self._staging[key] = {"next": 0, "slots": [
                (torch.empty(a.shape, dtype=a.dtype).pin_memory(), torch.empty(a.shape, dtype=a.dtype, device=self.device), torch.cuda.Event())
                for _ in range(2)]}
        host, dev, done = ring["slots"][ring["next"]]
        ring["next"] ^= 1
        done.synchronize()
        host.copy_(a)
        dev.copy_(host, non_blocking=True)
        done.record()
        return dev

    def _to_device(self, a, allowed, name):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        if not isinstance(a, torch.Tensor):
            raise TypeError("TSDFVolume.integrate: %s must be a numpy array or a tensor" % name)
        if a.dtype not in allowed:
            a = a.to(allowed[-1])
        if not a.is_cuda:
            return self._upload(a, name)
        if a.device != self.device:
            raise RuntimeError("TSDFVolume.integrate: %s lives on %s, the volume on %s" % (name, a.device, self.device))
        return a.contiguous()

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
        """Integrate an RGB-D frame (:220-323).  color_im (H,W,3) RGB uint8 / float; depth_im (H,W) float32 / float64, 0 =
        invalid; cam_intr (3,3); cam_pose (4,4) camera-to-world; obs_weight: weight of this observation."""
        intr = np.asarray(cam_intr.cpu() if isinstance(cam_intr, torch.Tensor) else cam_intr).astype(np.float32)      # :197
        pose = np.asarray(cam_pose.cpu() if isinstance(cam_pose, torch.Tensor) else cam_pose)
        inv = np.ascontiguousarray(np.linalg.inv(pose), dtype=np.float64)              # :285, host logic as in the reference
        intr4 = (ctypes.c_float * 4)(float(intr[0, 0]), float(intr[1, 1]), float(intr[0, 2]), float(intr[1, 2]))
        inv16 = (ctypes.c_double * 16)(*[float(v) for v in inv.reshape(-1)])
        with torch.cuda.device(self.device):
            depth = self._to_device(depth_im, (torch.float64, torch.float32), "depth_im")
            color = self._to_device(color_im, (torch.uint8, torch.float32), "color_im")
            if depth.dim() != 2 or tuple(color.shape) != (depth.shape[0], depth.shape[1], 3):
                raise RuntimeError("TSDFVolume.integrate: expected depth (H,W) and colour (H,W,3), got %s and %s" % (tuple(depth.shape), tuple(color.shape)))
            N.check(N.lib().dvmvs_tsdf_integrate(
                self._tsdf_vol.data_ptr(), self._weight_vol.data_ptr(), self._color_vol.data_ptr(),
                int(self._vol_dim[0]), int(self._vol_dim[1]), int(self._vol_dim[2]), self._origin_c, self._voxel_size,
                self._trunc_margin, color.data_ptr(), 1 if color.dtype == torch.uint8 else 0, depth.data_ptr(),
                1 if depth.dtype == torch.float64 else 0, int(depth.shape[0]), int(depth.shape[1]), intr4, inv16,
                float(obs_weight), self._updated.data_ptr(), _stream()), "tsdf_integrate")

    # ---- read-back ----------------------------------------------------------------------------------------------------
    def get_volume(self):
        """(tsdf, colour) as numpy arrays, like the reference (:325-328; its GPU mode copies device -> host here too)."""
        return self._tsdf_vol.cpu().numpy(), self._color_vol.cpu().numpy()

    def get_volume_tensors(self):
        """(tsdf, weight, colour) CUDA tensors, no copy."""
        return self._tsdf_vol, self._weight_vol, self._color_vol

    def updated_voxels(self):
        """Total number of voxel updates so far (synchronises)."""
        return int(self._updated.item())

    def _colors_at(self, color_vol, verts_ind):
        rgb_vals = color_vol[verts_ind[:, 0], verts_ind[:, 1], verts_ind[:, 2]]
        b = np.floor(rgb_vals / self._color_const)
        g = np.floor((rgb_vals - b * self._color_const) / 256)
        r = rgb_vals - b * self._color_const - g * 256
        return np.floor(np.asarray([r, g, b])).T.astype(np.uint8)

    def get_mesh(self):
        """:344-358 -- marching cubes is scikit-image's in the reference as well."""
        try:
            from skimage import measure
        except ImportError:
            raise RuntimeError("TSDFVolume.get_mesh needs scikit-image (marching cubes), as the reference does")
        tsdf_vol, color_vol = self.get_volume()
        mc = getattr(measure, "marching_cubes_lewiner", None) or measure.marching_cubes
        verts, faces, norms, _ = mc(tsdf_vol, level=0)
        verts_ind = np.round(verts).astype(int)
        verts = verts * self._voxel_size + self._vol_origin
        return verts, faces, norms, self._colors_at(color_vol, verts_ind)

    def get_point_cloud(self):
        """:329-342."""
        verts, _, _, colors = self.get_mesh()
        return np.hstack([verts, colors])


class TSDFFusion(object):
    """Host-side helpers of the reference's driver (:360-480): frustum bounds and the per-frame loop."""

    @staticmethod
    def rigid_transform(xyz, transform):
        """(N,3) points through a 4x4 transform (:361-367)."""
        xyz = np.asarray(xyz)
        homogeneous = np.concatenate([xyz, np.ones((len(xyz), 1), dtype=np.float32)], axis=1)
        return np.dot(transform, homogeneous.T).T[:, :3]

    @staticmethod
    def get_view_frustum(depth_im, cam_intr, cam_pose):
        """The 5 corners (apex + far plane) of the camera frustum in world coordinates, (3,5) (:369-382)."""
        im_h, im_w = depth_im.shape[0], depth_im.shape[1]
        max_depth = float(np.max(depth_im))
        us = np.array([0, 0, 0, im_w, im_w], dtype=np.float64)
        vs = np.array([0, 0, im_h, 0, im_h], dtype=np.float64)
        zs = np.array([0, max_depth, max_depth, max_depth, max_depth])
        pts = np.stack([(us - cam_intr[0, 2]) * zs / cam_intr[0, 0], (vs - cam_intr[1, 2]) * zs / cam_intr[1, 1], zs], axis=1)
        return TSDFFusion.rigid_transform(pts, cam_pose).T

    @staticmethod
    def calculate_volume_bounds(depth_maps, poses, K):
        """:465-475 (bounds start at the origin, as in the reference)."""
        assert len(depth_maps) == len(poses)
        bounds = np.zeros((3, 2))
        for depth_map, pose in zip(depth_maps, poses):
            pts = TSDFFusion.get_view_frustum(depth_map, K, pose)
            bounds[:, 0] = np.minimum(bounds[:, 0], pts.min(axis=1))
            bounds[:, 1] = np.maximum(bounds[:, 1], pts.max(axis=1))
        return bounds

    @staticmethod
    def integrate(tsdf_volume, images, depths, poses, K, obs_weight=1.):
        """The fusion loop of :442-463 without the mesh files: every frame into the volume; returns frames per second
        (device-timed when on a GPU)."""
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for image, depth, pose in zip(images, depths, poses):
            tsdf_volume.integrate(image, depth, K, pose, obs_weight=obs_weight)
        stop.record()
        stop.synchronize()
        return len(images) / max(start.elapsed_time(stop) * 1e-3, 1e-9)
