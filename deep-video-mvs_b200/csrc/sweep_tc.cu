// Plane-sweep warp + correlate as "correlate the epipolar band on the tensor cores, then interpolate scalars".  sm_100a.
//
// Reference behaviour reproduced (paths relative to the reference root): dvmvs/utils.py:45-107
// (calculate_cost_volume_by_warping / cost_volume_fusion), dot-product mode, C = 32.
//
//   cost[p, d] = 1/(32 M) * sum_m sum_{t in 2x2 taps} w_t(p, d, m) * ( f1[p] . f2_m[q_t(p, d, m)] )
//
// The cost is LINEAR in the four taps, so the 32-channel dot products S[q][p] = f2[q] . f1[p] can be formed first and the
// bilinear blend applied to four SCALARS per sample.  Consecutive planes move a pixel's sampling position by well under a
// pixel along its epipolar line and neighbouring reference pixels sample neighbouring measurement pixels, so a tile of
// reference pixels needs S only on a narrow band of measurement pixels around the epipolar segment -- and every entry of
// that band is reused by ~20 of the tile's (pixel, plane) samples.  S over (band x tile) is a dense contraction with K = 32:
//
//   * CTA = a 16 x 4 tile of reference pixels (N = 64).  Its fp16 feature rows are ONE 4-D TMA box in SWIZZLE_64B layout
//     (the UMMA B operand).
//   * For a measurement frame the planes are processed in chunks.  For a chunk, one thread per plane maps the tile's four
//     corners through the plane's homography (utils.py:51-73); the bounding boxes are merged into per-row [xmin, xmax]
//     tables with shared-memory atomics: a sheared band that follows the epipolar line, not a bounding rectangle.  A chunk
//     that does not fit the band capacity is halved; a single plane that does not fit (or whose denominator changes sign
//     over the tile) takes a direct gather path, so every pose is handled.
//   * The band is fetched by TMA, one {32 ch, 8 px, 1 row} box per 512-byte swizzle atom, with out-of-image pixels
//     zero-filled by the TMA unit (= grid_sample's zero padding for free); 128 band pixels form the A operand of
//     tcgen05.mma (M = 128, N = 64, K = 32: two K steps; fp16 (hi, lo) pairs issue hi*hi + lo*hi + hi*lo for
//     fp32-equivalent products, `terms` = 1 plain fp16).  Accumulators: TMEM, 64 columns per 128 band pixels.
//   * TMEM lane = band pixel q, column = reference pixel p.  Each thread copies its lane's 64 values to shared memory as
//     row q of S[q][p] (pitch 68 floats: conflict-free 16-byte stores), which makes the per-sample look-ups
//     S[q_t][p] conflict-free as well (lane = p).
//   * Look-up phase: one thread per (pixel, plane): homography in registers, perspective divide, four LDS + four FMA.
//     No bounds tests: positions are clamped to [-1, w] x [-1, h], where every tap is either inside the zero-filled band
//     or has weight 0.  Costs accumulate over the M frames in shared memory and are written once, coalesced.
//
// Per sample this moves ~(20 band entries x 4 B written + 16 B read) through shared memory instead of 512 B (fp32 features)
// through the L1 gather path of plane_sweep_c32_kernel.
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "tc_ptx.cuh"

namespace dvmvs {

constexpr int kStTileW = 16, kStTileH = 4, kStPix = kStTileW * kStTileH;     // reference-pixel tile (UMMA N = 64)
constexpr int kStThreads = 256;
constexpr int kStPitch = 68;            // floats per S row (64 + 4: conflict-free STS.128 / LDS.32)
constexpr int kStRows = 64;             // band rows per chunk
constexpr int kStTmemCols = 256;        // 4 M-tiles x 64 columns
constexpr int kStMaxQ = 512;            // band pixels per chunk (4 M-tiles of 128)
constexpr int kStMaxMeas = 8;
constexpr int kStMaxPlanes = 128;
constexpr int kStAccPitch = kStPix + 1; // cost accumulators [D][65]

struct SweepTcParams {
  CUtensorMap ref_map[2];               // fp16 planes [B][h][w][32]: hi, lo; box {32, 16, 4, 1}
  CUtensorMap meas_map[kStMaxMeas][2];  // box {32, 8, 1, 1}
  const __half* meas_planes[kStMaxMeas][2];   // raw pointers for the direct (fallback) path
  const __half* ref_planes[2];
  const float* pose2[kStMaxMeas];
  const float* pose1;
  const float* K;
  float* out;                           // [B][h][w][D]
  int B, h, w, D, M;
  int tiles_x, tiles_y;
  int qcap;                             // band capacity in pixels (multiple of 8, <= kStMaxQ)
  double inv_base, inv_step;
};

// same algebra as geometry.cu sweep_matrices (utils.py:51-56): G = K R K^-1, Kt = K t
__device__ __forceinline__ void st_matrices(const float* pose1, const float* pose2, const float* K, float* G, float* Kt) {
  float inv2[16], E[16];
  mat4_rigid_free_inverse(pose2, inv2);
  mat4_mul(inv2, pose1, E);
  float R[9] = {E[0], E[1], E[2], E[4], E[5], E[6], E[8], E[9], E[10]};
  float t[3] = {E[3], E[7], E[11]};
  float Kinv[9], KR[9];
  mat3_inverse(K, Kinv);
  mat3_mul(K, R, KR);
  mat3_mul(KR, Kinv, G);
#pragma unroll
  for (int i = 0; i < 3; ++i) Kt[i] = fmaf(K[i * 3 + 2], t[2], fmaf(K[i * 3 + 1], t[1], K[i * 3 + 0] * t[0]));
}

struct StSmem {          // fixed-size bookkeeping behind the big arrays
  float G[kStMaxMeas][12];
  int row_xmin[kStRows], row_xmax[kStRows], row_q[kStRows];       // row_q: q index of pixel x on this row = row_q + x
  int ymin, ymax, bad, fit, total_q, nrows;
  unsigned long long bar_ref, bar_band, bar_mma;
  uint32_t tmem_slot;
};

// sample position of pixel (uf, vf) on plane kd: same op sequence as plane_sweep_c32_kernel (<= 3 ulp from the reference)
__device__ __forceinline__ void st_position(const float* G, float4 kd, float uf, float vf, float sx, float sy, float wf, float hf,
                                            float& xs, float& ys, float& den) {
  const float q0 = fmaf(G[0], uf, fmaf(G[1], vf, G[2])) + kd.x;
  const float q1 = fmaf(G[3], uf, fmaf(G[4], vf, G[5])) + kd.y;
  const float q2 = fmaf(G[6], uf, fmaf(G[7], vf, G[8])) + kd.z;
  den = q2 + 1e-8f;
  const float r = __frcp_rn(den);
  xs = fminf(fmaxf(q0 * r * sx, -1.f), wf);        // NaN / -Inf -> -1, +Inf -> w: every tap outside the image or weight 0
  ys = fminf(fmaxf(q1 * r * sy, -1.f), hf);
}

// direct 2x2-tap gather of one sample from the fp16 planes (fallback path; same in-image semantics as the band path)
template <int TERMS>
__device__ __forceinline__ float st_direct_sample(const SweepTcParams& p, int m, int b, const float* f1, float xs, float ys) {
  const float x0f = floorf(xs), y0f = floorf(ys);
  const float fx = xs - x0f, fy = ys - y0f, gx = (x0f + 1.f) - xs, gy = (y0f + 1.f) - ys;
  const int x0 = (int)x0f, y0 = (int)y0f;
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int x = x0 + (t & 1), y = y0 + (t >> 1);
    if (x < 0 || x >= p.w || y < 0 || y >= p.h) continue;
    const float wt = ((t & 1) ? fx : gx) * ((t >> 1) ? fy : gy);
    const size_t off = (((size_t)b * p.h + y) * p.w + x) * 32;
    float dot = 0.f;
    for (int c = 0; c < 32; ++c) {
      float v = __half2float(p.meas_planes[m][0][off + c]);
      if (TERMS == 3) v += __half2float(p.meas_planes[m][1][off + c]);
      dot = fmaf(f1[c], v, dot);
    }
    acc = fmaf(dot, wt, acc);
  }
  return acc;
}

template <int TERMS>
__global__ void __launch_bounds__(kStThreads, 1) plane_sweep_tc_kernel(const __grid_constant__ SweepTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  // layout: [S / band (aliased)] [ref tile hi, lo] [acc] [kd] [bookkeeping]
  const uint32_t s_bytes = (uint32_t)p.qcap * kStPitch * 4u;
  float* S = reinterpret_cast<float*>(base_ptr);
  const uint32_t band_addr = base;                                   // band hi at +0, lo at +qcap*64 (dead once the MMAs completed)
  const uint32_t ref_off = (s_bytes + 1023u) & ~1023u;
  const uint32_t ref_addr = base + ref_off;                          // 64 rows x 64 B per plane
  float* acc = reinterpret_cast<float*>(base_ptr + ref_off + 2 * kStPix * 64);
  float4* s_kd = reinterpret_cast<float4*>(acc + p.D * kStAccPitch + 3);             // [D]
  s_kd = reinterpret_cast<float4*>(((uintptr_t)s_kd + 15) & ~(uintptr_t)15);
  StSmem* sm = reinterpret_cast<StSmem*>(s_kd + p.D);
  sm = reinterpret_cast<StSmem*>(((uintptr_t)sm + 15) & ~(uintptr_t)15);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int b = blockIdx.x / tiles_per_img;
  const int t_in = blockIdx.x - b * tiles_per_img;
  const int v0 = (t_in / p.tiles_x) * kStTileH, u0 = (t_in % p.tiles_x) * kStTileW;
  const int tw = min(kStTileW, p.w - u0), th = min(kStTileH, p.h - v0);      // valid extent of this tile

  const uint32_t bar_ref = smem_u32(&sm->bar_ref), bar_band = smem_u32(&sm->bar_band), bar_mma = smem_u32(&sm->bar_mma);
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.ref_map[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.meas_map[0][0]) : "memory");
    mbar_init(bar_ref, 1);
    mbar_init(bar_band, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_slot)), "n"(kStTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_slot;
  pdl_wait();

  // ---- reference tile (B operand) by TMA; pose algebra of all frames meanwhile
  if (tid == 0) {
    mbar_expect_tx(bar_ref, (TERMS == 3 ? 2u : 1u) * kStPix * 64u);
    tma_load_4d(ref_addr, &p.ref_map[0], bar_ref, 0, u0, v0, b);
    if (TERMS == 3) tma_load_4d(ref_addr + kStPix * 64, &p.ref_map[1], bar_ref, 0, u0, v0, b);
  }
  if (tid < p.M) {
    float G[9], Kt[3];
    st_matrices(p.pose1 + b * 16, p.pose2[tid] + b * 16, p.K + b * 9, G, Kt);
#pragma unroll
    for (int i = 0; i < 9; ++i) sm->G[tid][i] = G[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) sm->G[tid][9 + i] = Kt[i];
  }
  __syncthreads();

  const float sx = (float)(p.w - 1) / (float)p.w, sy = (float)(p.h - 1) / (float)p.h;      // align_corners "shrink" (App. A.1)
  const float wf = (float)p.w, hf = (float)p.h;
  // look-up role of this thread: pixel pl (lane within a 32-pixel half), planes d0 + (warp >> 1), +4, ...
  const int pl = (warp & 1) * 32 + lane;
  const int pty = pl >> 4, ptx = pl & 15;
  const bool pix_valid = (ptx < tw) && (pty < th);
  const float uf = (float)(u0 + min(ptx, tw - 1)), vf = (float)(v0 + min(pty, th - 1));
  const uint32_t hi_word = umma_hi_word(512u, 4u);                   // SBO = 8 rows x 64 B, SWIZZLE_64B
  const uint32_t idesc = (1u << 4) | ((uint32_t)(kStPix >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // D=f32, A=B=f16, K-major, N=64, M=128
  uint32_t band_phase = 0, mma_phase = 0;
  bool ref_ready = false;

  for (int m = 0; m < p.M; ++m) {
    const float* G = sm->G[m];
    for (int i = tid; i < p.D; i += kStThreads) {
      const float this_depth = (float)(1.0 / (p.inv_base + i * p.inv_step));       // utils.py:66
      s_kd[i] = make_float4(G[9] / this_depth, G[10] / this_depth, G[11] / this_depth, 0.f);      // utils.py:68
    }
    __syncthreads();
    // first guess of the chunk length from the displacement of the tile centre between the first and the last plane
    int dc;
    {
      float xa, ya, xb, yb, den;
      const float uc = (float)u0 + 0.5f * (float)(tw - 1), vc = (float)v0 + 0.5f * (float)(th - 1);
      st_position(G, s_kd[0], uc, vc, sx, sy, wf, hf, xa, ya, den);
      st_position(G, s_kd[p.D - 1], uc, vc, sx, sy, wf, hf, xb, yb, den);
      const float tile_q = (float)((kStTileW + 10) * (kStTileH + 3));          // +8-pixel row padding, +2 footprint, +1 slack
      const float per_plane = (fabsf(xb - xa) * (kStTileH + 3) + fabsf(yb - ya) * (kStTileW + 10)) / (float)(p.D - 1);
      float n = ((float)p.qcap * 0.9f - tile_q) / fmaxf(per_plane, 1e-3f);
      n = fminf(fmaxf(n, 1.f), (float)p.D);
      dc = (n == n) ? (int)n : 1;
    }
    float b0 = fmaf(G[0], uf, fmaf(G[1], vf, G[2]));
    float b1 = fmaf(G[3], uf, fmaf(G[4], vf, G[5]));
    float b2 = fmaf(G[6], uf, fmaf(G[7], vf, G[8]));

    int d0 = 0;
    while (d0 < p.D) {
      int nd = min(dc, p.D - d0);
      // ---------------- (a) band geometry of planes [d0, d0 + nd)
      if (tid < kStRows) { sm->row_xmin[tid] = INT_MAX; sm->row_xmax[tid] = INT_MIN; }
      if (tid == 0) { sm->ymin = INT_MAX; sm->ymax = INT_MIN; sm->bad = 0; }
      __syncthreads();
      int xl = 0, xh = -1, yl = 0, yh = -1;
      if (tid < nd) {
        const float4 kd = s_kd[d0 + tid];
        float xmn = 3.0e38f, xmx = -3.0e38f, ymn = 3.0e38f, ymx = -3.0e38f;
        int pos = 0, neg = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float cu = (float)(u0 + ((c & 1) ? tw - 1 : 0)), cv = (float)(v0 + ((c >> 1) ? th - 1 : 0));
          float xs, ys, den;
          st_position(G, kd, cu, cv, sx, sy, wf, hf, xs, ys, den);
          pos += (den > 1e-6f);
          neg += (den < -1e-6f);
          xmn = fminf(xmn, xs); xmx = fmaxf(xmx, xs); ymn = fminf(ymn, ys); ymx = fmaxf(ymx, ys);
        }
        if (pos != 4 && neg != 4) atomicOr(&sm->bad, 1);        // denominator vanishes / changes sign on the tile: not convex
        xl = (int)floorf(xmn - 1e-3f); xh = (int)floorf(xmx + 1e-3f) + 1;
        yl = (int)floorf(ymn - 1e-3f); yh = (int)floorf(ymx + 1e-3f) + 1;
        atomicMin(&sm->ymin, yl);
        atomicMax(&sm->ymax, yh);
      }
      __syncthreads();
      const int ymin = sm->ymin;
      if (tid < nd) {
        for (int y = yl; y <= yh; ++y) {
          const int r = y - ymin;
          if (r < kStRows) { atomicMin(&sm->row_xmin[r], xl); atomicMax(&sm->row_xmax[r], xh); }
        }
      }
      __syncthreads();
      if (warp == 0) {
        // rows -> 8-pixel groups -> q offsets (two rows per lane, exclusive scan over 64 rows)
        const int nrows = sm->ymax - ymin + 1;
        int g0 = 0, g1 = 0;
        const int r0 = 2 * lane, r1 = 2 * lane + 1;
        if (r0 < nrows && sm->row_xmax[r0] >= sm->row_xmin[r0]) g0 = (sm->row_xmax[r0] - sm->row_xmin[r0] + 8) >> 3;
        if (r1 < nrows && r1 < kStRows && sm->row_xmax[r1] >= sm->row_xmin[r1]) g1 = (sm->row_xmax[r1] - sm->row_xmin[r1] + 8) >> 3;
        int incl = g0 + g1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const int excl = incl - (g0 + g1);
        const int total = __shfl_sync(0xffffffffu, incl, 31) * 8;
        if (g0) sm->row_q[r0] = excl * 8 - sm->row_xmin[r0];
        if (g1) sm->row_q[r1] = (excl + g0) * 8 - sm->row_xmin[r1];
        if (lane == 0) {
          sm->nrows = nrows;
          sm->total_q = total;
          sm->fit = (nrows <= kStRows && total <= p.qcap && total > 0 && !sm->bad) ? 1 : 0;
        }
      }
      __syncthreads();
      const int fit = sm->fit, bad = sm->bad, total_q = sm->total_q, nrows = sm->nrows;
      __syncthreads();                       // everybody holds the verdict before the tables are reset for the next attempt
      if (!fit) {
        if (nd > 1 && !bad) {                // halve the chunk and retry
          dc = max(1, nd >> 1);
          continue;
        }
        // ---------------- direct path for these planes (band does not fit, or the homography is degenerate on the tile)
        if (!ref_ready) { mbar_wait(bar_ref, 0); ref_ready = true; }
        {
          const int pp = tid & (kStPix - 1);
          const int ty = pp >> 4, tx = pp & 15;
          if (tx < tw && ty < th) {
            float f1[32];
            const size_t roff = (((size_t)b * p.h + v0 + ty) * p.w + u0 + tx) * 32;
            for (int c = 0; c < 32; ++c) {
              f1[c] = __half2float(p.ref_planes[0][roff + c]);
              if (TERMS == 3) f1[c] += __half2float(p.ref_planes[1][roff + c]);
            }
            for (int d = d0 + (tid >> 6); d < d0 + nd; d += kStThreads / kStPix) {
              float xs, ys, den;
              const float4 kd = s_kd[d];
              const float q0 = fmaf(G[0], (float)(u0 + tx), fmaf(G[1], (float)(v0 + ty), G[2])) + kd.x;
              const float q1 = fmaf(G[3], (float)(u0 + tx), fmaf(G[4], (float)(v0 + ty), G[5])) + kd.y;
              const float q2 = fmaf(G[6], (float)(u0 + tx), fmaf(G[7], (float)(v0 + ty), G[8])) + kd.z;
              den = q2 + 1e-8f;
              const float r = __frcp_rn(den);
              xs = q0 * r * sx; ys = q1 * r * sy;
              float val = 0.f;
              if (xs > -1.f && xs < wf && ys > -1.f && ys < hf) val = st_direct_sample<TERMS>(p, m, b, f1, xs, ys);
              float* a = acc + d * kStAccPitch + pp;
              *a = (m == 0) ? val * (1.f / 32.f) : fmaf(val, 1.f / 32.f, *a);
            }
          }
        }
        __syncthreads();
        d0 += nd;
        continue;
      }
      // ---------------- (b) band -> shared memory by TMA (warp 0: one row per lane, one box per 8-pixel group)
      if (warp == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic accesses of the S region precede these async writes
        if (lane == 0) mbar_expect_tx(bar_band, (uint32_t)total_q * 64u * (TERMS == 3 ? 2u : 1u));
        __syncwarp();
        for (int r = lane; r < nrows; r += 32) {
          const int xmin = sm->row_xmin[r], xmax = sm->row_xmax[r];
          if (xmax < xmin) continue;
          const int ng = (xmax - xmin + 8) >> 3;
          const int q0 = sm->row_q[r] + xmin;
          for (int g = 0; g < ng; ++g) {
            const uint32_t dst = band_addr + (uint32_t)(q0 + 8 * g) * 64u;
            tma_load_4d(dst, &p.meas_map[m][0], bar_band, 0, xmin + 8 * g, ymin + r, b);
            if (TERMS == 3) tma_load_4d(dst + (uint32_t)p.qcap * 64u, &p.meas_map[m][1], bar_band, 0, xmin + 8 * g, ymin + r, b);
          }
        }
      } else if (warp == 1 && lane == 0) {
        // ---------------- (c) S = band . tile^T on the tensor cores
        if (!ref_ready) mbar_wait(bar_ref, 0);
        mbar_wait(bar_band, band_phase);
        tc_fence_after();
        const int n_mt = (total_q + 127) >> 7;
        const uint32_t b_hi = umma_lo_word(ref_addr, 16), b_lo = umma_lo_word(ref_addr + kStPix * 64, 16);
        for (int mt = 0; mt < n_mt; ++mt) {
          const uint32_t a_hi = umma_lo_word(band_addr + (uint32_t)mt * 8192u, 16);
          const uint32_t a_lo = umma_lo_word(band_addr + (uint32_t)p.qcap * 64u + (uint32_t)mt * 8192u, 16);
          const uint32_t d_tmem = tmem_base + (uint32_t)mt * kStPix;
#pragma unroll
          for (int k = 0; k < 2; ++k) tc_mma_f16_words(d_tmem, a_hi + 2 * k, hi_word, b_hi + 2 * k, hi_word, idesc, k > 0);
          if (TERMS == 3) {
#pragma unroll
            for (int k = 0; k < 2; ++k) tc_mma_f16_words(d_tmem, a_lo + 2 * k, hi_word, b_hi + 2 * k, hi_word, idesc, 1u);
#pragma unroll
            for (int k = 0; k < 2; ++k) tc_mma_f16_words(d_tmem, a_hi + 2 * k, hi_word, b_lo + 2 * k, hi_word, idesc, 1u);
          }
        }
        tc_commit(bar_mma);
      }
      ref_ready = true;
      band_phase ^= 1u;
      // ---------------- (d) TMEM -> S[q][p] in shared memory (the band bytes are dead once the MMAs have completed)
      mbar_wait(bar_mma, mma_phase);
      mma_phase ^= 1u;
      tc_fence_after();
      {
        const int n_mt = (total_q + 127) >> 7;
        const int wq = warp & 3;
        for (int mt = warp >> 2; mt < n_mt; mt += 2) {
          const int q = mt * 128 + wq * 32 + lane;
          float vals[64];
          tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * kStPix), vals);
          tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * kStPix + 32), vals + 32);
          if (q < p.qcap) {
            float4* dst = reinterpret_cast<float4*>(S + (size_t)q * kStPitch);
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j] = make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      // ---------------- (e) look-ups: thread = (pixel, plane), four scalars per sample
      if (pix_valid) {
        const int qmax = p.qcap - 2;
        for (int d = d0 + (warp >> 1); d < d0 + nd; d += 4) {
          const float4 kd = s_kd[d];
          const float q0 = b0 + kd.x, q1 = b1 + kd.y, q2 = b2 + kd.z;
          const float r = __frcp_rn(q2 + 1e-8f);
          const float xs = fminf(fmaxf(q0 * r * sx, -1.f), wf), ys = fminf(fmaxf(q1 * r * sy, -1.f), hf);
          const float x0f = floorf(xs), y0f = floorf(ys);
          const float fx = xs - x0f, fy = ys - y0f, gx = (x0f + 1.f) - xs, gy = (y0f + 1.f) - ys;
          const int ix = (int)x0f;
          const int iy = min(max((int)y0f - ymin, 0), nrows - 2);
          const int qa = min(max(sm->row_q[iy] + ix, 0), qmax), qb = min(max(sm->row_q[iy + 1] + ix, 0), qmax);
          const float* sa = S + qa * kStPitch + pl;
          const float* sb = S + qb * kStPitch + pl;
          const float val = fmaf(sb[kStPitch], fx * fy, fmaf(sb[0], gx * fy, fmaf(sa[kStPitch], fx * gy, sa[0] * (gx * gy))));
          float* a = acc + d * kStAccPitch + pl;
          *a = (m == 0) ? val * (1.f / 32.f) : fmaf(val, 1.f / 32.f, *a);          // utils.py:82 (/C), summed over frames (:102)
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // S was accessed through the generic proxy; the next band arrives by TMA
      __syncthreads();
      d0 += nd;
      dc = nd;          // keep the chunk length that worked
      if (dc < p.D && total_q * 2 <= p.qcap) dc = min(p.D, dc * 2);      // and grow it again when the band got narrow
    }
  }
  // ---- coalesced write-out: rows of the tile are contiguous [tw][D] spans of the channel-last cost volume
  for (int ty = 0; ty < th; ++ty) {
    float* o = p.out + (((size_t)b * p.h + v0 + ty) * p.w + u0) * p.D;
    for (int i = tid; i < tw * p.D; i += kStThreads) {
      const int tx = i / p.D, d = i - tx * p.D;
      o[i] = acc[d * kStAccPitch + ty * kStTileW + tx] / (float)p.M;      // utils.py:105-106
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kStTmemCols) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------- host side
// Encoded tensor maps are cached per (pointer, shape, box): cuTensorMapEncodeTiled costs a few microseconds per map and a
// sweep needs 2 (M + 1) of them; inside CUDA graphs the cost disappears, on the eager module path it is most of the call.
struct MapKey {
  const void* ptr;
  int B, h, w, bw, bh;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && B == o.B && h == o.h && w == o.w && bw == o.bw && bh == o.bh; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t x = (size_t)k.ptr;
    x ^= ((size_t)k.B * 0x9E3779B97F4A7C15ull) ^ ((size_t)k.h << 20) ^ ((size_t)k.w << 36) ^ ((size_t)k.bw << 52) ^ ((size_t)k.bh << 58);
    return x;
  }
};
static std::mutex g_map_mutex;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

static int feature_map(CUtensorMap* out, const void* ptr, int B, int h, int w, int box_w, int box_h) {
  MapKey key{ptr, B, h, w, box_w, box_h};
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) { *out = it->second; return DVMVS_OK; }
  cuuint64_t dims[4] = {32, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
  cuuint64_t strides[3] = {64, (cuuint64_t)w * 64, (cuuint64_t)h * w * 64};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = tensor_map_encoder()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("plane_sweep_tc: cuTensorMapEncodeTiled(B=%d h=%d w=%d box=%dx%d) failed: %d", B, h, w, box_w, box_h, (int)r);
    return DVMVS_EINVAL;
  }
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache.emplace(key, *out);
  return DVMVS_OK;
}

}  // namespace dvmvs

using namespace dvmvs;

extern "C" int dvmvs_plane_sweep_tc(const void* ref_hi, const void* ref_lo, const void* const* meas_hi_host, const void* const* meas_lo_host,
                                    const float* pose1, const float* const* pose2_host, const float* K, float* cost_out, int B, int h, int w,
                                    int D, int M, float min_depth, float max_depth, int terms, dvmvs_stream_t stream) {
  DVMVS_REQUIRE(ref_hi && meas_hi_host && pose1 && pose2_host && K && cost_out, "plane_sweep_tc: null pointer");
  DVMVS_REQUIRE(tensor_map_encoder() != nullptr, "plane_sweep_tc: cuTensorMapEncodeTiled entry point not available");
  DVMVS_REQUIRE(B > 0 && h > 1 && w > 1, "plane_sweep_tc: bad shape B=%d h=%d w=%d", B, h, w);
  DVMVS_REQUIRE(D >= 2 && D <= kStMaxPlanes, "plane_sweep_tc: D=%d outside [2,%d]", D, kStMaxPlanes);
  DVMVS_REQUIRE(M >= 1 && M <= kStMaxMeas, "plane_sweep_tc: M=%d outside [1,%d]", M, kStMaxMeas);
  DVMVS_REQUIRE(terms == 1 || terms == 3, "plane_sweep_tc: terms=%d", terms);
  DVMVS_REQUIRE(terms == 1 || (ref_lo && meas_lo_host), "plane_sweep_tc: terms=3 needs the lo planes");
  DVMVS_REQUIRE(min_depth > 0.f && max_depth > min_depth, "plane_sweep_tc: bad depth range");
  DVMVS_REQUIRE((uintptr_t)ref_hi % 16 == 0 && (uintptr_t)ref_lo % 16 == 0, "plane_sweep_tc: reference planes not 16-byte aligned");
  static SweepTcParams p;        // large (tensor maps): filled under the lock below, passed by value at launch
  static std::mutex launch_mutex;
  std::lock_guard<std::mutex> lock(launch_mutex);
  memset(&p, 0, sizeof(p));
  p.ref_planes[0] = (const __half*)ref_hi;
  p.ref_planes[1] = (const __half*)ref_lo;
  int rc = feature_map(&p.ref_map[0], p.ref_planes[0], B, h, w, kStTileW, kStTileH);
  if (rc != DVMVS_OK) return rc;
  if (terms == 3) {
    rc = feature_map(&p.ref_map[1], p.ref_planes[1], B, h, w, kStTileW, kStTileH);
    if (rc != DVMVS_OK) return rc;
  }
  for (int m = 0; m < M; ++m) {
    DVMVS_REQUIRE(meas_hi_host[m] && pose2_host[m] && (uintptr_t)meas_hi_host[m] % 16 == 0, "plane_sweep_tc: bad measurement pointer %d", m);
    p.meas_planes[m][0] = (const __half*)meas_hi_host[m];
    rc = feature_map(&p.meas_map[m][0], p.meas_planes[m][0], B, h, w, 8, 1);
    if (rc != DVMVS_OK) return rc;
    if (terms == 3) {
      DVMVS_REQUIRE(meas_lo_host[m] && (uintptr_t)meas_lo_host[m] % 16 == 0, "plane_sweep_tc: bad measurement lo pointer %d", m);
      p.meas_planes[m][1] = (const __half*)meas_lo_host[m];
      rc = feature_map(&p.meas_map[m][1], p.meas_planes[m][1], B, h, w, 8, 1);
      if (rc != DVMVS_OK) return rc;
    }
    p.pose2[m] = pose2_host[m];
  }
  p.pose1 = pose1; p.K = K; p.out = cost_out;
  p.B = B; p.h = h; p.w = w; p.D = D; p.M = M;
  p.tiles_x = (w + kStTileW - 1) / kStTileW;
  p.tiles_y = (h + kStTileH - 1) / kStTileH;
  p.inv_base = 1.0 / (double)max_depth;                                   // utils.py:59-60
  p.inv_step = (1.0 / (double)min_depth - 1.0 / (double)max_depth) / (double)(D - 1);
  static const int qcap_env = []() { const char* e = getenv("DVMVS_SWEEP_QCAP"); return e ? atoi(e) : 0; }();
  int qcap = kStMaxQ;
  if (terms == 3) qcap = 448;                     // hi + lo bands of 512 pixels would not leave room beside S
  if (qcap_env >= 64 && qcap_env <= kStMaxQ) qcap = qcap_env & ~7;
  p.qcap = qcap;
  const size_t smem = 1024 + (((size_t)qcap * kStPitch * 4 + 1023) & ~(size_t)1023) + 2 * kStPix * 64 + (size_t)D * kStAccPitch * 4 + 32 +
                      (size_t)D * 16 + 32 + sizeof(StSmem);
  DVMVS_REQUIRE(smem <= 227 * 1024, "plane_sweep_tc: shared memory %zu too large (D=%d)", smem, D);
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(plane_sweep_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(plane_sweep_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set[dev] = true;
  }
  const int ctas = B * p.tiles_x * p.tiles_y;
  if (terms == 3) launch_k(plane_sweep_tc_kernel<3>, dim3(ctas), dim3(kStThreads), smem, (cudaStream_t)stream, p);
  else launch_k(plane_sweep_tc_kernel<1>, dim3(ctas), dim3(kStThreads), smem, (cudaStream_t)stream, p);
  return check_launch("plane_sweep_tc_kernel");
}
