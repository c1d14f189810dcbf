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
//   * Prologue, all threads, all frames at once: the per-plane homography terms (utils.py:51-68) and, per plane, the
//     bounding box of the tile's four corners mapped through that plane's homography.
//   * Warp-specialised pipeline over CHUNKS of consecutive planes of one measurement frame:
//       producer warp   plans the chunk (lanes own the 64 circular band-row slots, slot = y & 63; per-row [xmin, xmax] =
//                       union of the boxes of the chunk's planes -- a sheared band that follows the epipolar line, not a
//                       bounding rectangle; a chunk that exceeds the band capacity is halved, a single plane that does not
//                       fit or whose denominator changes sign over the tile goes to a direct gather path, so every pose is
//                       handled), fetches the band rows by TMA (one {32 ch, 32 px, 1 row} box per 2 KB, out-of-image pixels
//                       zero-filled by the TMA unit = grid_sample's zero padding for free) and issues tcgen05.mma: 128 band
//                       pixels are the A operand (M = 128, N = 64, K = 32: two K steps; fp16 (hi, lo) pairs issue
//                       hi*hi + lo*hi + hi*lo for fp32-equivalent products, `terms` = 1 plain fp16); accumulators in TMEM,
//                       64 columns per 128 band pixels;
//       16 consumer warps  TMEM lane = band pixel q, column = tile pixel p: each thread copies its lane's 64 values to shared
//                       memory as row q of S[q][p] (fp16 pre-scaled by 1/32 at terms = 1, fp32 at terms = 3; row pitches
//                       that make the 16-byte stores conflict-free), then one thread per (pixel, plane): homography in
//                       registers, perspective divide, four LDS + four FMA.  No bounds tests: positions are clamped to
//                       [-1, w] x [-1, h], where every tap is inside the zero-filled band or has weight 0.
//     The producer runs one chunk ahead (band buffer, accumulators and chunk descriptors are handed back and forth with
//     mbarriers), so planning, TMA latency and the MMAs of chunk i+1 hide behind the look-ups of chunk i.
//   * Costs accumulate over the M frames in shared memory and are written once, coalesced.
//
// Per sample this moves ~(20 band entries x 2-4 B written + 8-16 B read) through shared memory instead of 512 B (fp32
// features) through the L1 gather path of plane_sweep_c32_kernel.
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "tc_ptx.cuh"

namespace dvmvs {

constexpr int kStTileW = 16, kStTileH = 4, kStPix = kStTileW * kStTileH;     // reference-pixel tile (UMMA N = 64)
constexpr int kStConsumers = 512;       // warps 0..15: TMEM -> shared copy + look-ups
constexpr int kStThreads = kStConsumers + 32;     // warp 16: planner + TMA producer + MMA issuer
constexpr int kStRows = 64;             // band rows per chunk (circular slots: slot = y & 63)
constexpr int kStBox = 32;              // band pixels per TMA box (one box = 32 rows of 64 B = four SWIZZLE_64B atoms)
constexpr int kStMaxMeas = 8;
constexpr int kStMaxPlanes = 128;
constexpr int kStMaxMD = 512;           // M * D (homography tables in shared memory)
constexpr int kStAccPitch = kStPix + 1; // cost accumulators [D][65]

template <int TERMS>
struct StCfg {
  // S[q][p]: TERMS == 1 -> fp16 (pre-scaled by 1/32), pitch 72 halves; TERMS == 3 -> fp32, pitch 68 floats.  Both pitches make
  // the 16-byte row stores of eight consecutive q conflict-free.
  static constexpr int kSPitchBytes = (TERMS == 1) ? 144 : 272;
  static constexpr int kBandBytesPerQ = (TERMS == 1) ? 64 : 128;      // hi (+ lo) feature rows
  static constexpr int kMaxQ = (TERMS == 1) ? 1024 : 512;             // 8 / 4 M-tiles of 128 band pixels
  static constexpr int kTmemCols = (TERMS == 1) ? 512 : 256;          // 64 columns per M-tile
};

struct SweepTcParams {
  CUtensorMap ref_map[2];               // fp16 planes [B][h][w][32]: hi, lo; box {32, 16, 4, 1}
  CUtensorMap meas_map[kStMaxMeas][2];  // box {32, 32, 1, 1}
  const __half* meas_planes[kStMaxMeas][2];   // raw pointers for the direct (fallback) path
  const __half* ref_planes[2];
  const float* pose2[kStMaxMeas];
  const float* pose1;
  const float* K;
  float* out;                           // [B][h][w][D]
  int B, h, w, D, M;
  int tiles_x, tiles_y;
  int qcap;                             // band capacity in pixels (multiple of 32)
  float depth[kStMaxPlanes];            // plane depths, computed on the host in double like the reference (utils.py:59-66)
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

constexpr int kStMaxChunks = 48;
struct StChunk {         // one chunk of consecutive planes of one frame, planned in the prologue
  short m, d0, nd, band, total_q, ylo;
  short row_q[kStRows];            // q index of pixel x on the band row in slot (y & 63) = row_q + x
  short xmn[kStRows];              // first band pixel of the row
  unsigned char nb[kStRows];       // 32-pixel boxes of the row
};

struct StSmem {          // fixed-size bookkeeping behind the big arrays
  float G[kStMaxMeas][12];
  StChunk chunk[kStMaxChunks];
  int frame_n[kStMaxMeas], frame_fail[kStMaxMeas];
  int n_chunks, any_fail;
  unsigned long long bar_ref, bar_band, bar_mma, bar_tmem_empty;
  uint32_t tmem_slot;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kStConsumers) : "memory"); }

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

// estimated band pixels of a chunk of n planes: rows x 32-pixel boxes per row, from the displacement (dx, dy) of the tile
// centre per plane.  Only a first guess -- the planner verifies the real band and halves the chunk when it does not fit.
__device__ __forceinline__ float st_band_estimate(float n, float dx, float dy) {
  const float rows = (float)(kStTileH + 3) + dy * n;
  const float planes_per_row = fminf(n, (float)(kStTileH + 3) / fmaxf(dy, 1e-6f));
  const float width = (float)(kStTileW + 3) + dx * planes_per_row;
  return rows * (float)kStBox * ceilf(width * (1.f / kStBox));
}

template <int TERMS>
__global__ void __launch_bounds__(kStThreads, 1) plane_sweep_tc_kernel(const __grid_constant__ SweepTcParams p) {
  using Cfg = StCfg<TERMS>;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  // layout: [band hi (+ lo)] [ref tile hi (+ lo)] [S] [acc] [kd] [plane boxes] [bookkeeping]
  const uint32_t band_addr = base;                                               // hi at +0, lo at +qcap*64
  const uint32_t band_bytes = (uint32_t)p.qcap * Cfg::kBandBytesPerQ;            // multiple of 2048
  const uint32_t ref_addr = base + band_bytes;                                   // 64 rows x 64 B per plane
  uint8_t* S = base_ptr + band_bytes + (TERMS == 3 ? 2 : 1) * kStPix * 64;
  float* acc = reinterpret_cast<float*>(S + (size_t)p.qcap * Cfg::kSPitchBytes);
  float4* s_kd = reinterpret_cast<float4*>(acc + ((p.D * kStAccPitch + 3) & ~3));      // [M][D]
  int4* s_box = reinterpret_cast<int4*>(s_kd + p.M * p.D);                              // [M][D] {xl, xh, yl, yh}; xl > xh: degenerate
  StSmem* sm = reinterpret_cast<StSmem*>(s_box + p.M * p.D);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool producer = warp == kStConsumers / 32;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int b = blockIdx.x / tiles_per_img;
  const int t_in = blockIdx.x - b * tiles_per_img;
  const int v0 = (t_in / p.tiles_x) * kStTileH, u0 = (t_in % p.tiles_x) * kStTileW;
  const int tw = min(kStTileW, p.w - u0), th = min(kStTileH, p.h - v0);      // valid extent of this tile

  const uint32_t bar_ref = smem_u32(&sm->bar_ref), bar_band = smem_u32(&sm->bar_band), bar_mma = smem_u32(&sm->bar_mma);
  const uint32_t bar_tmem_empty = smem_u32(&sm->bar_tmem_empty);
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.ref_map[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.meas_map[0][0]) : "memory");
    mbar_init(bar_ref, 1);
    mbar_init(bar_band, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_tmem_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (producer) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_slot)), "n"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_slot;
  pdl_wait();

  // ---- reference tile (B operand) by TMA; pose algebra and per-plane geometry of ALL frames meanwhile (one pass, all threads)
  if (producer && lane == 0) {
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
  for (int i = tid; i < p.M * p.D; i += kStThreads) {
    const int m = i / p.D, d = i - m * p.D;
    const float* G = sm->G[m];
    const float this_depth = p.depth[d];                                                     // utils.py:66
    const float4 kd = make_float4(G[9] / this_depth, G[10] / this_depth, G[11] / this_depth, 0.f);      // utils.py:68
    s_kd[i] = kd;
    // bounding box of the tile's image on this plane: a homography with a denominator of one sign maps the (convex) tile
    // onto a convex quadrilateral, so the four corners bound every sample; +-1e-3 px absorbs fp32 rounding
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
    int4 bb;
    bb.x = (int)floorf(xmn - 1e-3f); bb.y = (int)floorf(xmx + 1e-3f) + 1;
    bb.z = (int)floorf(ymn - 1e-3f); bb.w = (int)floorf(ymx + 1e-3f) + 1;
    if ((pos != 4 && neg != 4) || bb.w - bb.z >= kStRows) { bb.x = 1; bb.y = 0; }       // degenerate on this tile: direct path
    s_box[i] = bb;
  }
  if (tid < p.M) {
    // first guess of the planes per chunk of frame tid from the motion of the tile centre between the first and the last plane
    const float* G = sm->G[tid];
    const float uc = (float)u0 + 0.5f * (float)(tw - 1), vc = (float)v0 + 0.5f * (float)(th - 1);
    const float da = p.depth[0], db = p.depth[p.D - 1];
    float xa, ya, xb, yb, den;
    st_position(G, make_float4(G[9] / da, G[10] / da, G[11] / da, 0.f), uc, vc, sx, sy, wf, hf, xa, ya, den);
    st_position(G, make_float4(G[9] / db, G[10] / db, G[11] / db, 0.f), uc, vc, sx, sy, wf, hf, xb, yb, den);
    const float dx = fabsf(xb - xa) / (float)(p.D - 1), dy = fabsf(yb - ya) / (float)(p.D - 1);
    // uniform chunks: the smallest number of chunks whose estimated band fits, within this frame's share of the chunk list
    const int budget = max(1, kStMaxChunks / p.M);
    int nch = 1;
    while (nch < budget && st_band_estimate((float)((p.D + nch - 1) / nch), dx, dy) * 1.05f > (float)p.qcap) ++nch;
    sm->frame_n[tid] = (p.D + nch - 1) / nch;
    sm->frame_fail[tid] = 0;
  }
  __syncthreads();
  // ---- chunk plan: uniform chunks per frame, every chunk verified against the real per-plane boxes by one warp (lanes own
  // the 64 circular row slots); a frame with a chunk that does not fit is re-planned with shorter chunks as long as its
  // share of the chunk list allows -- what still does not fit then takes the direct path
  const int min_n = (p.D + max(1, kStMaxChunks / p.M) - 1) / max(1, kStMaxChunks / p.M);
  for (int round = 0; round < 10; ++round) {
    if (tid == 0) {
      int cnt = 0;
      for (int m = 0; m < p.M; ++m) {
        const int n = sm->frame_n[m];
        for (int d0 = 0; d0 < p.D && cnt < kStMaxChunks; d0 += n) {
          StChunk* ch = &sm->chunk[cnt++];
          ch->m = (short)m; ch->d0 = (short)d0; ch->nd = (short)min(n, p.D - d0);
        }
      }
      sm->n_chunks = cnt;
      sm->any_fail = 0;
    }
    __syncthreads();
    for (int k = warp; k < sm->n_chunks; k += kStThreads / 32) {
      StChunk* ch = &sm->chunk[k];
      const int m = ch->m, d0 = ch->d0, nd = ch->nd;
      int xmn0 = INT_MAX, xmx0 = INT_MIN, xmn1 = INT_MAX, xmx1 = INT_MIN, ylo = INT_MAX, yhi = INT_MIN, degenerate = 0;
#pragma unroll 4
      for (int d = d0; d < d0 + nd; ++d) {
        const int4 bb = s_box[m * p.D + d];
        degenerate |= (bb.x > bb.y);
        ylo = min(ylo, bb.z); yhi = max(yhi, bb.w);
        const int span = bb.w - bb.z;
        if (((2 * lane - bb.z) & (kStRows - 1)) <= span) { xmn0 = min(xmn0, bb.x); xmx0 = max(xmx0, bb.y); }
        if (((2 * lane + 1 - bb.z) & (kStRows - 1)) <= span) { xmn1 = min(xmn1, bb.x); xmx1 = max(xmx1, bb.y); }
      }
      const int n0 = (!degenerate && xmx0 >= xmn0) ? (xmx0 - xmn0 + kStBox) / kStBox : 0;
      const int n1 = (!degenerate && xmx1 >= xmn1) ? (xmx1 - xmn1 + kStBox) / kStBox : 0;
      int incl = n0 + n1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int total_q = __shfl_sync(0xffffffffu, incl, 31) * kStBox;
      const int q0 = (incl - n0 - n1) * kStBox, q1 = q0 + n0 * kStBox;
      const bool fits = !degenerate && yhi - ylo + 1 <= kStRows && total_q <= p.qcap && total_q > 0;
      if (fits) {
        ch->row_q[2 * lane] = (short)(n0 ? q0 - xmn0 : 0);
        ch->row_q[2 * lane + 1] = (short)(n1 ? q1 - xmn1 : 0);
        ch->xmn[2 * lane] = (short)(n0 ? xmn0 : 0);
        ch->xmn[2 * lane + 1] = (short)(n1 ? xmn1 : 0);
        ch->nb[2 * lane] = (unsigned char)n0;
        ch->nb[2 * lane + 1] = (unsigned char)n1;
      }
      if (lane == 0) {
        ch->band = fits ? 1 : 0;
        ch->total_q = (short)(fits ? total_q : 0);
        ch->ylo = (short)(fits ? ylo : 0);
        if (!fits && !degenerate && sm->frame_n[m] > min_n && nd > 1) { sm->frame_fail[m] = 1; sm->any_fail = 1; }      // shorter chunks may fit
      }
    }
    __syncthreads();
    if (!sm->any_fail) break;
    __syncthreads();
    if (tid < p.M && sm->frame_fail[tid]) {
      sm->frame_n[tid] = max(min_n, (sm->frame_n[tid] * 3) >> 2);
      sm->frame_fail[tid] = 0;
    }
    __syncthreads();
  }
  const int n_chunks = sm->n_chunks;

  if (producer) {
    // =============================== TMA producer + MMA issuer (one warp), one chunk ahead of the consumers ===============================
    const uint32_t hi_word = umma_hi_word(512u, 4u);                   // SBO = 8 rows x 64 B, SWIZZLE_64B
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kStPix >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f32 += f16 x f16, K-major, N=64, M=128
    const uint32_t b_hi = umma_lo_word(ref_addr, 16), b_lo = umma_lo_word(ref_addr + kStPix * 64, 16);
    int n_band = 0;
    for (int k = 0; k < n_chunks; ++k) {
      const StChunk* ch = &sm->chunk[k];
      if (ch->band != 1) continue;                                     // direct chunk: nothing to stage
      const int m = ch->m, total_q = ch->total_q, ylo = ch->ylo;
      // ---- band rows -> shared memory by TMA (the previous band chunk's MMAs must have consumed the band buffer)
      if (n_band > 0) mbar_wait(bar_mma, (uint32_t)((n_band - 1) & 1));
      if (lane == 0) mbar_expect_tx(bar_band, (uint32_t)total_q * (uint32_t)Cfg::kBandBytesPerQ);
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int slot_r = 2 * lane + e;
        const int nb = ch->nb[slot_r], xmn = ch->xmn[slot_r], qs = ch->row_q[slot_r] + xmn;
        const int y = ylo + ((slot_r - ylo) & (kStRows - 1));
        for (int j = 0; j < nb; ++j) {
          const uint32_t dst = band_addr + (uint32_t)(qs + kStBox * j) * 64u;
          tma_load_4d(dst, &p.meas_map[m][0], bar_band, 0, xmn + kStBox * j, y, b);
          if (TERMS == 3) tma_load_4d(dst + (uint32_t)p.qcap * 64u, &p.meas_map[m][1], bar_band, 0, xmn + kStBox * j, y, b);
        }
      }
      if (lane == 0) {
        // ---- S = band . tile^T on the tensor cores
        if (n_band == 0) mbar_wait(bar_ref, 0);
        mbar_wait(bar_band, (uint32_t)(n_band & 1));
        if (n_band > 0) mbar_wait(bar_tmem_empty, (uint32_t)((n_band - 1) & 1));      // consumers drained the previous accumulators
        tc_fence_after();
        const int n_mt = (total_q + 127) >> 7;
        for (int mt = 0; mt < n_mt; ++mt) {
          const uint32_t a_hi = umma_lo_word(band_addr + (uint32_t)mt * 8192u, 16);
          const uint32_t a_lo = umma_lo_word(band_addr + (uint32_t)p.qcap * 64u + (uint32_t)mt * 8192u, 16);
          const uint32_t d_tmem = tmem_base + (uint32_t)mt * kStPix;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) tc_mma_f16_words(d_tmem, a_hi + 2 * kk, hi_word, b_hi + 2 * kk, hi_word, idesc, kk > 0);
          if (TERMS == 3) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) tc_mma_f16_words(d_tmem, a_lo + 2 * kk, hi_word, b_hi + 2 * kk, hi_word, idesc, 1u);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) tc_mma_f16_words(d_tmem, a_hi + 2 * kk, hi_word, b_lo + 2 * kk, hi_word, idesc, 1u);
          }
        }
        tc_commit(bar_mma);
      }
      ++n_band;
      __syncwarp();
    }
  } else {
    // =============================== consumers: TMEM -> S[q][p], then one thread per (pixel, plane) ===============================
    const int pl = tid & (kStPix - 1);                 // pixel of this thread; planes d0 + (tid >> 6), + 8, ...
    const int pty = pl >> 4, ptx = pl & 15;
    const bool pix_valid = (ptx < tw) && (pty < th);
    const float uf = (float)(u0 + min(ptx, tw - 1)), vf = (float)(v0 + min(pty, th - 1));
    const int wq = warp & 3;
    int n_band = 0, cur_m = -1;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    bool ref_ready = false;
    for (int k = 0; k < n_chunks; ++k) {
      const StChunk* ch = &sm->chunk[k];
      const int m = ch->m, d0 = ch->d0, nd = ch->nd, is_band = (ch->band == 1), total_q = ch->total_q;
      const float* G = sm->G[m];
      if (m != cur_m) {
        cur_m = m;
        b0 = fmaf(G[0], uf, fmaf(G[1], vf, G[2]));
        b1 = fmaf(G[3], uf, fmaf(G[4], vf, G[5]));
        b2 = fmaf(G[6], uf, fmaf(G[7], vf, G[8]));
      }
      const float4* kdm = s_kd + m * p.D;
      if (is_band) {
        mbar_wait(bar_mma, (uint32_t)(n_band & 1));
        ++n_band;
        tc_fence_after();
        // ---- TMEM lane = band pixel q, column = tile pixel p  ->  row q of S (all look-ups of the previous chunk are done:
        // consumer barrier at the end of the loop body)
        const int n_mt = (total_q + 127) >> 7;
        for (int mt = warp >> 2; mt < n_mt; mt += kStConsumers / 128) {
          const int q = mt * 128 + wq * 32 + lane;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float vals[32];
            tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * kStPix + half * 32), vals);
            if (q < total_q) {
              if (TERMS == 1) {
                uint4* dst = reinterpret_cast<uint4*>(S + (size_t)q * Cfg::kSPitchBytes + half * 64);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  __half2 h[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(vals[8 * j + 2 * e] * (1.f / 32.f), vals[8 * j + 2 * e + 1] * (1.f / 32.f));
                  dst[j] = *reinterpret_cast<const uint4*>(h);
                }
              } else {
                float4* dst = reinterpret_cast<float4*>(S + (size_t)q * Cfg::kSPitchBytes + half * 128);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
              }
            }
          }
        }
        tc_fence_before();
        consumer_barrier();
        if (tid == 0) mbar_arrive(bar_tmem_empty);         // the producer may overwrite the accumulators
        // ---- look-ups: four scalars per sample, no bounds tests (clamped positions, zero-filled band)
        if (pix_valid) {
          const short* row_q = ch->row_q;
          const int qmax = total_q - 2;
#pragma unroll 2
          for (int d = d0 + (tid >> 6); d < d0 + nd; d += kStConsumers / kStPix) {
            const float4 kd = kdm[d];
            const float q0 = b0 + kd.x, q1 = b1 + kd.y, q2 = b2 + kd.z;
            const float r = __frcp_rn(q2 + 1e-8f);
            const float xs = fminf(fmaxf(q0 * r * sx, -1.f), wf), ys = fminf(fmaxf(q1 * r * sy, -1.f), hf);
            const float x0f = floorf(xs), y0f = floorf(ys);
            const float fx = xs - x0f, fy = ys - y0f, gx = (x0f + 1.f) - xs, gy = (y0f + 1.f) - ys;
            const int ix = (int)x0f, iy = (int)y0f;
            const int qa = min(max(row_q[iy & (kStRows - 1)] + ix, 0), qmax), qb = min(max(row_q[(iy + 1) & (kStRows - 1)] + ix, 0), qmax);
            float s00, s01, s10, s11;
            if (TERMS == 1) {
              const __half* sa = reinterpret_cast<const __half*>(S + (size_t)qa * Cfg::kSPitchBytes) + pl;
              const __half* sb = reinterpret_cast<const __half*>(S + (size_t)qb * Cfg::kSPitchBytes) + pl;
              s00 = __half2float(sa[0]); s01 = __half2float(sa[Cfg::kSPitchBytes / 2]);
              s10 = __half2float(sb[0]); s11 = __half2float(sb[Cfg::kSPitchBytes / 2]);
            } else {
              const float* sa = reinterpret_cast<const float*>(S + (size_t)qa * Cfg::kSPitchBytes) + pl;
              const float* sb = reinterpret_cast<const float*>(S + (size_t)qb * Cfg::kSPitchBytes) + pl;
              s00 = sa[0]; s01 = sa[Cfg::kSPitchBytes / 4]; s10 = sb[0]; s11 = sb[Cfg::kSPitchBytes / 4];
            }
            float val = fmaf(s11, fx * fy, fmaf(s10, gx * fy, fmaf(s01, fx * gy, s00 * (gx * gy))));
            if (TERMS == 3) val *= (1.f / 32.f);                          // utils.py:82 (/C); the fp16 S is stored pre-scaled
            float* a = acc + d * kStAccPitch + pl;
            *a = (m == 0) ? val : *a + val;                               // summed over the measurement frames (utils.py:102)
          }
        }
      } else {
        // ---- direct path for these planes (band does not fit, or the homography is degenerate on the tile)
        if (!ref_ready) { mbar_wait(bar_ref, 0); ref_ready = true; }
        if (pix_valid) {
          float f1[32];
          const size_t roff = (((size_t)b * p.h + v0 + pty) * p.w + u0 + ptx) * 32;
          for (int c = 0; c < 32; ++c) {
            f1[c] = __half2float(p.ref_planes[0][roff + c]);
            if (TERMS == 3) f1[c] += __half2float(p.ref_planes[1][roff + c]);
          }
          for (int d = d0 + (tid >> 6); d < d0 + nd; d += kStConsumers / kStPix) {
            const float4 kd = kdm[d];
            const float q0 = b0 + kd.x, q1 = b1 + kd.y, q2 = b2 + kd.z;
            const float r = __frcp_rn(q2 + 1e-8f);
            const float xs = q0 * r * sx, ys = q1 * r * sy;
            float val = 0.f;
            if (xs > -1.f && xs < wf && ys > -1.f && ys < hf) val = st_direct_sample<TERMS>(p, m, b, f1, xs, ys) * (1.f / 32.f);
            float* a = acc + d * kStAccPitch + pl;
            *a = (m == 0) ? val : *a + val;
          }
        }
      }
      consumer_barrier();                                  // S is free again
    }
    // ---- coalesced write-out: rows of the tile are contiguous [tw][D] spans of the channel-last cost volume
    for (int ty = 0; ty < th; ++ty) {
      float* o = p.out + (((size_t)b * p.h + v0 + ty) * p.w + u0) * p.D;
      for (int i = tid; i < tw * p.D; i += kStConsumers) {
        const int tx = i / p.D, d = i - tx * p.D;
        o[i] = acc[d * kStAccPitch + ty * kStTileW + tx] / (float)p.M;      // utils.py:105-106
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (producer) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
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
    rc = feature_map(&p.meas_map[m][0], p.meas_planes[m][0], B, h, w, kStBox, 1);
    if (rc != DVMVS_OK) return rc;
    if (terms == 3) {
      DVMVS_REQUIRE(meas_lo_host[m] && (uintptr_t)meas_lo_host[m] % 16 == 0, "plane_sweep_tc: bad measurement lo pointer %d", m);
      p.meas_planes[m][1] = (const __half*)meas_lo_host[m];
      rc = feature_map(&p.meas_map[m][1], p.meas_planes[m][1], B, h, w, kStBox, 1);
      if (rc != DVMVS_OK) return rc;
    }
    p.pose2[m] = pose2_host[m];
  }
  p.pose1 = pose1; p.K = K; p.out = cost_out;
  p.B = B; p.h = h; p.w = w; p.D = D; p.M = M;
  p.tiles_x = (w + kStTileW - 1) / kStTileW;
  p.tiles_y = (h + kStTileH - 1) / kStTileH;
  {
    const double inv_base = 1.0 / (double)max_depth;                      // utils.py:59-60
    const double inv_step = (1.0 / (double)min_depth - 1.0 / (double)max_depth) / (double)(D - 1);
    for (int d = 0; d < D; ++d) p.depth[d] = (float)(1.0 / (inv_base + d * inv_step));     // utils.py:66 (double, then fp32 divide)
  }
  DVMVS_REQUIRE(M * D <= kStMaxMD, "plane_sweep_tc: M * D = %d exceeds %d", M * D, kStMaxMD);
  // band capacity: whatever shared memory is left after the fixed arrays, in whole 32-pixel boxes, capped by the TMEM columns
  static const int qcap_env = []() { const char* e = getenv("DVMVS_SWEEP_QCAP"); return e ? atoi(e) : 0; }();
  const size_t fixed = 1024 + (size_t)(terms == 3 ? 2 : 1) * kStPix * 64 + (size_t)((D * kStAccPitch + 3) & ~3) * 4 + (size_t)M * D * 32 + sizeof(StSmem) + 64;
  const int per_q = (terms == 3) ? StCfg<3>::kSPitchBytes + StCfg<3>::kBandBytesPerQ : StCfg<1>::kSPitchBytes + StCfg<1>::kBandBytesPerQ;
  int qcap = (int)((227 * 1024 - fixed) / per_q) & ~(kStBox - 1);
  qcap = min(qcap, terms == 3 ? StCfg<3>::kMaxQ : StCfg<1>::kMaxQ);
  if (qcap_env >= kStBox && qcap_env <= qcap) qcap = qcap_env & ~(kStBox - 1);
  DVMVS_REQUIRE(qcap >= 2 * kStBox, "plane_sweep_tc: no shared memory left for the band (D=%d, M=%d)", D, M);
  p.qcap = qcap;
  const size_t smem = fixed + (size_t)qcap * per_q;
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
