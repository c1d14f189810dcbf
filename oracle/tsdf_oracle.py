"""TEST INFRASTRUCTURE ONLY -- never imported by the product (deep-video-mvs_b200/).  CPU restatement (numpy, no numba) of the
reference's TSDF integration, SURVEY section 8 row f4: `TSDFVolume` of sample-data/run-tsdf-reconstruction.py, CPU mode
(use_gpu=False), the path the reference runs wherever pycuda is absent.  Pinned bit-for-bit against goldens produced by the
unmodified script (oracle/make_golden_tsdf.py -> tests/golden/tsdf.npz; tests/test_tsdf.py).

The reference mixes precisions; each step below names the dtype the reference computes in (numba's scalar typing rules for the
@njit helpers, NumPy-2 promotion for the vectorised part):
  * world point  = float32( float64(origin32) + voxel_size(float64) * float64(coord) )                      (:181-191 vox2world)
  * camera point = float64: inv(cam_pose) (float64, LAPACK) times [x y z 1] (float32 promoted)              (:285, :360-366)
  * pixel        = int( round_half_even( X * float64(fx32) / Z + float64(cx32) ) ), float64                  (:193-204 cam2pix)
  * valid        = pixel inside the image and Z > 0; depth > 0; depth - Z >= -trunc  (float64)               (:289-300)
  * dist         = min(1, (depth - Z) / trunc), float64                                                      (:301)
  * w_new        = float32( float64(w_old32) + obs_weight )                                                  (:214)
  * tsdf         = float32( ( float64( float32(w_old32 * tsdf32) ) + obs_weight * dist ) / float64(w_new32) ) (:215)
  * colours      : all float32 (obs_weight as float32), np.round = half-even, min(255, .), fold b*65536+g*256+r (:311-323)
"""
import numpy as np


class TSDFVolume(object):
    def __init__(self, vol_bnds, voxel_size):
        # run-tsdf-reconstruction.py:34-66
        vol_bnds = np.array(vol_bnds, dtype=np.float64)
        assert vol_bnds.shape == (3, 2)
        self.voxel_size = float(voxel_size)
        self.trunc_margin = 5 * self.voxel_size
        self.vol_dim = np.ceil((vol_bnds[:, 1] - vol_bnds[:, 0]) / self.voxel_size).astype(int)
        self.vol_origin = vol_bnds[:, 0].astype(np.float32)
        self.tsdf = np.ones(self.vol_dim, dtype=np.float32)
        self.weight = np.zeros(self.vol_dim, dtype=np.float32)
        self.color = np.zeros(self.vol_dim, dtype=np.float32)
        # :163-176 voxel grid coordinates, x-major (C order)
        xv, yv, zv = np.meshgrid(range(self.vol_dim[0]), range(self.vol_dim[1]), range(self.vol_dim[2]), indexing="ij")
        self.vox_coords = np.stack([xv.reshape(-1), yv.reshape(-1), zv.reshape(-1)], axis=1).astype(int)

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.0):
        # :220-323
        im_h, im_w = depth_im.shape
        obs_weight = float(obs_weight)
        c = np.asarray(color_im).astype(np.float32)
        folded = np.floor(c[..., 2] * np.float32(65536) + c[..., 1] * np.float32(256) + c[..., 0])        # float32 (:235-236)

        coords = self.vox_coords
        world = (self.vol_origin.astype(np.float64)[None, :] + self.voxel_size * coords.astype(np.float32).astype(np.float64)).astype(np.float32)
        xyz_h = np.hstack([world, np.ones((len(world), 1), dtype=np.float32)])
        cam = np.dot(np.linalg.inv(cam_pose), xyz_h.T).T[:, :3]                                          # float64
        intr = np.asarray(cam_intr).astype(np.float32)
        fx, fy, cx, cy = [np.float64(v) for v in (intr[0, 0], intr[1, 1], intr[0, 2], intr[1, 2])]
        with np.errstate(all="ignore"):
            px = np.round(cam[:, 0] * fx / cam[:, 2] + cx)
            py = np.round(cam[:, 1] * fy / cam[:, 2] + cy)
        pz = cam[:, 2]
        valid_pix = (px >= 0) & (px < im_w) & (py >= 0) & (py < im_h) & (pz > 0)                           # NaN compares false
        ix = np.where(valid_pix, px, 0).astype(np.int64)
        iy = np.where(valid_pix, py, 0).astype(np.int64)
        depth_val = np.zeros(len(px))
        depth_val[valid_pix] = depth_im[iy[valid_pix], ix[valid_pix]]
        depth_diff = depth_val - pz
        valid = (depth_val > 0) & (depth_diff >= -self.trunc_margin)
        dist = np.minimum(1, depth_diff / self.trunc_margin)

        vx, vy, vz = coords[valid, 0], coords[valid, 1], coords[valid, 2]
        w_old = self.weight[vx, vy, vz]
        tsdf_old = self.tsdf[vx, vy, vz]
        with np.errstate(all="ignore"):
            w_new = (w_old.astype(np.float64) + obs_weight).astype(np.float32)
            tsdf_new = (((w_old * tsdf_old).astype(np.float64) + obs_weight * dist[valid]) / w_new.astype(np.float64)).astype(np.float32)
        self.weight[vx, vy, vz] = w_new
        self.tsdf[vx, vy, vz] = tsdf_new

        f32 = np.float32
        ow = f32(obs_weight)
        old = self.color[vx, vy, vz]
        old_b = np.floor(old / f32(65536))
        old_g = np.floor((old - old_b * f32(65536)) / f32(256))
        old_r = old - old_b * f32(65536) - old_g * f32(256)
        new = folded[iy[valid], ix[valid]]
        new_b = np.floor(new / f32(65536))
        new_g = np.floor((new - new_b * f32(65536)) / f32(256))
        new_r = new - new_b * f32(65536) - new_g * f32(256)
        with np.errstate(all="ignore"):
            new_b = np.minimum(f32(255), np.round((w_old * old_b + ow * new_b) / w_new))
            new_g = np.minimum(f32(255), np.round((w_old * old_g + ow * new_g) / w_new))
            new_r = np.minimum(f32(255), np.round((w_old * old_r + ow * new_r) / w_new))
        self.color[vx, vy, vz] = new_b * f32(65536) + new_g * f32(256) + new_r
        return int(valid.sum())

    def get_volume(self):
        return self.tsdf, self.color
