// TSDF integration of one posed RGB-D frame into a voxel volume -- SURVEY.md section 8 row f4, the consumer of the depth
// maps this library predicts: `TSDFVolume.integrate` of the reference's sample-data/run-tsdf-reconstruction.py:220-323.
// The reference has two implementations, an inline pycuda kernel in float32 (:80-152) and a numba / numpy CPU path in mixed
// float32 / float64 (:181-218, :283-323) that it falls back to wherever pycuda is absent (the case in this image).  This
// kernel follows the CPU path operation by operation, in the precision each operation has there, so its volumes equal the
// reference's bit for bit (tests/test_tsdf.py against goldens of the unmodified script):
//   world  = f32( f64(origin) + voxel_size * f64(coord) )                                     vox2world
//   cam    = inv(cam_pose) * [world 1] in f64; per row an FMA chain in k order starting from the rounded first product --
//            what the 4 x N double GEMM of np.dot evaluates                                   rigid_transform
//   pixel  = rint( X * f64(fx) / Z + f64(cx) ) (half-even), valid iff inside the image and Z > 0    cam2pix
//   tsdf   = f32( ( f64( f32(w_old * tsdf_old) ) + obs * dist ) / f64(w_new) ),  w_new = f32( f64(w_old) + obs )
//   colour : float32 throughout, rintf, min(255, .) with NaN propagation as np.minimum
// No multiply-add contraction anywhere a rounding would be skipped (explicit _rn intrinsics).
//
// Culling: 98 % of a room-sized volume is outside the frustum or the truncation band of any one frame, and the float64
// projection (two double divisions) is what the kernel would spend its time on.  Each voxel is therefore first projected in
// float32 (9 FMAs) and dropped when it is PROVABLY behind the camera or more than a pixel outside the image: the test uses
// error bounds of the float32 evaluation computed on the host from the volume extent (eps), so a voxel the float64 path would
// update is never dropped (tests compare whole volumes with the reference bit for bit).  Survivors take the exact path.
//
// HBM-bound: one thread per voxel, z fastest (the reference's C-order [x][y][z]) so a warp reads / writes 128 B rows of each
// of the three volumes; voxels outside the frustum or the truncation band touch no volume memory at all.  Algorithmic bytes:
// 24 B per UPDATED voxel (tsdf, weight, colour: read + write fp32) + the frame (4 B depth + 3..12 B colour per pixel, L2
// resident).  The inverse pose, float32 intrinsics and trunc margin arrive by value in the launch parameters.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace dvmvs {

struct TsdfParams {
  float* tsdf;
  float* weight;
  float* color;
  const void* color_im;   // [h][w][3] RGB, uint8 or float32
  const void* depth_im;   // [h][w] float32 or float64
  int dim_x, dim_y, dim_z, im_h, im_w;
  float origin[3];
  double voxel_size, trunc, obs;
  double T[12];           // rows 0..2 of inv(cam_pose)
  double fx, fy, cx, cy;  // float32 intrinsics widened
  unsigned long long* updated;   // optional counter of updated voxels (nullptr = none)
  float Tf[12];           // float32 copy of T for the conservative pre-test
  float eps[3];           // bound on |float32 camera coordinate - float64 camera coordinate| per axis over the volume
  float fxf, fyf, lo_x, hi_x, lo_y, hi_y;   // lo = cx + 1.5, hi = im_w + 1.5 - cx (one pixel of margin on each side)
  int cull;
};

__device__ __forceinline__ float np_minimum(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
__device__ __forceinline__ double np_minimum(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

__device__ __forceinline__ void unfold(float c, float& b, float& g, float& r) {
  b = floorf(__fdiv_rn(c, 65536.f));
  const float rest = __fsub_rn(c, __fmul_rn(b, 65536.f));
  g = floorf(__fdiv_rn(rest, 256.f));
  r = __fsub_rn(rest, __fmul_rn(g, 256.f));
}

template <typename ColorT, typename DepthT>
__global__ void __launch_bounds__(256) tsdf_integrate_kernel(TsdfParams p) {
  pdl_launch_dependents();
  const long long n = (long long)p.dim_x * p.dim_y * p.dim_z;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false;
  if (idx < n) {
    int x, y, z;
    if (n <= 0xffffffffLL) {       // 32-bit index arithmetic (64-bit division is emulated: it was most of this kernel's time)
      const unsigned i = (unsigned)idx, xy = i / (unsigned)p.dim_z;
      z = (int)(i - xy * (unsigned)p.dim_z);
      x = (int)(xy / (unsigned)p.dim_y);
      y = (int)(xy - (unsigned)x * (unsigned)p.dim_y);
    } else {
      z = (int)(idx % p.dim_z);
      const long long xy = idx / p.dim_z;
      y = (int)(xy % p.dim_y);
      x = (int)(xy / p.dim_y);
    }
    const float wxf = __double2float_rn(__dadd_rn((double)p.origin[0], __dmul_rn(p.voxel_size, (double)x)));
    const float wyf = __double2float_rn(__dadd_rn((double)p.origin[1], __dmul_rn(p.voxel_size, (double)y)));
    const float wzf = __double2float_rn(__dadd_rn((double)p.origin[2], __dmul_rn(p.voxel_size, (double)z)));
    bool candidate = true;
    if (p.cull) {
      const float X = fmaf(p.Tf[0], wxf, fmaf(p.Tf[1], wyf, fmaf(p.Tf[2], wzf, p.Tf[3])));
      const float Y = fmaf(p.Tf[4], wxf, fmaf(p.Tf[5], wyf, fmaf(p.Tf[6], wzf, p.Tf[7])));
      const float Z = fmaf(p.Tf[8], wxf, fmaf(p.Tf[9], wyf, fmaf(p.Tf[10], wzf, p.Tf[11])));
      if (Z + p.eps[2] <= 0.f) candidate = false;                 // certainly behind the camera
      else if (Z - p.eps[2] > 0.f) {                               // certainly in front: test the image borders with slack
        const float ux = X * p.fxf, uy = Y * p.fyf;
        const float sx = p.fxf * p.eps[0] + (fabsf(p.lo_x) + fabsf(p.hi_x)) * p.eps[2] + 1e-6f * (fabsf(ux) + (fabsf(p.lo_x) + fabsf(p.hi_x)) * Z);
        const float sy = p.fyf * p.eps[1] + (fabsf(p.lo_y) + fabsf(p.hi_y)) * p.eps[2] + 1e-6f * (fabsf(uy) + (fabsf(p.lo_y) + fabsf(p.hi_y)) * Z);
        // pixel < -1.5  <=>  X fx + (cx + 1.5) Z < 0 ;   pixel > w + 0.5  <=>  X fx - (w + 1.5 - cx) Z > 0  (with a pixel to spare)
        if (ux + p.lo_x * Z < -sx || ux - p.hi_x * Z > sx || uy + p.lo_y * Z < -sy || uy - p.hi_y * Z > sy) candidate = false;
      }
    }
    const double wx = (double)wxf, wy = (double)wyf, wz = (double)wzf;
    if (candidate) {
    const double cxp = __fma_rn(p.T[3], 1.0, __fma_rn(p.T[2], wz, __fma_rn(p.T[1], wy, __dmul_rn(p.T[0], wx))));
    const double cyp = __fma_rn(p.T[7], 1.0, __fma_rn(p.T[6], wz, __fma_rn(p.T[5], wy, __dmul_rn(p.T[4], wx))));
    const double czp = __fma_rn(p.T[11], 1.0, __fma_rn(p.T[10], wz, __fma_rn(p.T[9], wy, __dmul_rn(p.T[8], wx))));
    const double px = rint(__dadd_rn(__ddiv_rn(__dmul_rn(cxp, p.fx), czp), p.cx));
    const double py = rint(__dadd_rn(__ddiv_rn(__dmul_rn(cyp, p.fy), czp), p.cy));
    if (czp > 0.0 && px >= 0.0 && px < (double)p.im_w && py >= 0.0 && py < (double)p.im_h) {
      pdl_wait();
      const size_t pix = (size_t)py * p.im_w + (size_t)px;
      const double depth = (double)((const DepthT*)p.depth_im)[pix];
      const double diff = __dsub_rn(depth, czp);
      if (depth > 0.0 && diff >= -p.trunc) {
        hit = true;
        const double dist = np_minimum(1.0, __ddiv_rn(diff, p.trunc));
        const float w_old = p.weight[idx], t_old = p.tsdf[idx], c_old = p.color[idx];
        const float w_new = __double2float_rn(__dadd_rn((double)w_old, p.obs));
        const double num = __dadd_rn((double)__fmul_rn(w_old, t_old), __dmul_rn(p.obs, dist));
        p.weight[idx] = w_new;
        p.tsdf[idx] = __double2float_rn(__ddiv_rn(num, (double)w_new));
        const ColorT* c = (const ColorT*)p.color_im + pix * 3;
        const float folded = floorf(__fadd_rn(__fadd_rn(__fmul_rn((float)c[2], 65536.f), __fmul_rn((float)c[1], 256.f)), (float)c[0]));
        float ob, og, orr, nb, ng, nr;
        unfold(c_old, ob, og, orr);
        unfold(folded, nb, ng, nr);
        const float ow = (float)p.obs;
        nb = np_minimum(255.f, rintf(__fdiv_rn(__fadd_rn(__fmul_rn(w_old, ob), __fmul_rn(ow, nb)), w_new)));
        ng = np_minimum(255.f, rintf(__fdiv_rn(__fadd_rn(__fmul_rn(w_old, og), __fmul_rn(ow, ng)), w_new)));
        nr = np_minimum(255.f, rintf(__fdiv_rn(__fadd_rn(__fmul_rn(w_old, orr), __fmul_rn(ow, nr)), w_new)));
        p.color[idx] = __fadd_rn(__fadd_rn(__fmul_rn(nb, 65536.f), __fmul_rn(ng, 256.f)), nr);
      }
    }
    }
  }
  if (p.updated) {
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(p.updated, (unsigned long long)__popc(m));
  }
}

}  // namespace dvmvs

using namespace dvmvs;

extern "C" int dvmvs_tsdf_integrate(float* tsdf_vol, float* weight_vol, float* color_vol, int dim_x, int dim_y, int dim_z,
                                    const float* vol_origin3, double voxel_size, double trunc_margin, const void* color_im,
                                    int color_is_u8, const void* depth_im, int depth_is_f64, int im_h, int im_w,
                                    const float* intr4, const double* world_to_cam16, double obs_weight,
                                    unsigned long long* updated_count, dvmvs_stream_t stream) {
  DVMVS_REQUIRE(tsdf_vol && weight_vol && color_vol && vol_origin3 && color_im && depth_im && intr4 && world_to_cam16, "tsdf_integrate: null argument");
  DVMVS_REQUIRE(dim_x > 0 && dim_y > 0 && dim_z > 0 && im_h > 0 && im_w > 0, "tsdf_integrate: bad extent %d x %d x %d, image %d x %d", dim_x, dim_y, dim_z, im_h, im_w);
  const long long n = (long long)dim_x * dim_y * dim_z;
  DVMVS_REQUIRE((n + 255) / 256 <= 0x7fffffffLL, "tsdf_integrate: volume of %lld voxels exceeds one launch", n);
  TsdfParams p;
  p.tsdf = tsdf_vol; p.weight = weight_vol; p.color = color_vol;
  p.color_im = color_im; p.depth_im = depth_im;
  p.dim_x = dim_x; p.dim_y = dim_y; p.dim_z = dim_z; p.im_h = im_h; p.im_w = im_w;
  for (int i = 0; i < 3; ++i) p.origin[i] = vol_origin3[i];
  p.voxel_size = voxel_size; p.trunc = trunc_margin; p.obs = obs_weight;
  for (int i = 0; i < 12; ++i) p.T[i] = world_to_cam16[i];
  p.fx = (double)intr4[0]; p.fy = (double)intr4[1]; p.cx = (double)intr4[2]; p.cy = (double)intr4[3];
  p.updated = updated_count;
  // conservative float32 pre-test: per camera axis r, |fl32 evaluation - exact| <= 2^-20 * (sum_k |T[r][k]| max|w_k| + |T[r][3]|)
  // (four roundings of 2^-24 each on the products / sums, plus 2^-24 relative on each float32 copy of T: 16x headroom)
  double maxabs[3];
  const int dims[3] = {dim_x, dim_y, dim_z};
  for (int k = 0; k < 3; ++k) {
    const double a = fabs((double)vol_origin3[k]), b = fabs((double)vol_origin3[k] + voxel_size * (double)dims[k]);
    maxabs[k] = a > b ? a : b;
  }
  bool finite = true;
  for (int r = 0; r < 3; ++r) {
    double sum = fabs(p.T[4 * r + 3]);
    for (int k = 0; k < 3; ++k) sum += fabs(p.T[4 * r + k]) * maxabs[k];
    p.eps[r] = (float)(sum * 9.5367431640625e-07) + 1e-30f;
    for (int k = 0; k < 4; ++k) {
      p.Tf[4 * r + k] = (float)p.T[4 * r + k];
      finite = finite && isfinite(p.Tf[4 * r + k]);
    }
    finite = finite && isfinite(p.eps[r]);
  }
  p.fxf = fabsf(intr4[0]); p.fyf = fabsf(intr4[1]);
  p.lo_x = intr4[2] + 1.5f; p.hi_x = (float)im_w + 1.5f - intr4[2];
  p.lo_y = intr4[3] + 1.5f; p.hi_y = (float)im_h + 1.5f - intr4[3];
  // the border inequalities assume positive focal lengths; anything unusual (negative / non-finite) takes the exact path only
  static const bool cull_env = []() { const char* e = getenv("DVMVS_TSDF_CULL"); return !(e && e[0] == '0'); }();
  p.cull = (cull_env && finite && intr4[0] > 0.f && intr4[1] > 0.f && isfinite(intr4[2]) && isfinite(intr4[3])) ? 1 : 0;
  const dim3 grid((unsigned)((n + 255) / 256)), block(256);
  cudaStream_t s = (cudaStream_t)stream;
  if (color_is_u8) {
    if (depth_is_f64) launch_k(tsdf_integrate_kernel<unsigned char, double>, grid, block, 0, s, p);
    else launch_k(tsdf_integrate_kernel<unsigned char, float>, grid, block, 0, s, p);
  } else {
    if (depth_is_f64) launch_k(tsdf_integrate_kernel<float, double>, grid, block, 0, s, p);
    else launch_k(tsdf_integrate_kernel<float, float>, grid, block, 0, s, p);
  }
  return check_launch("tsdf_integrate_kernel");
}
