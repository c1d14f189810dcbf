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

#include "tc_ptx.cuh"

namespace dvmvs {

constexpr int kStTileW = 16, kStTileH = 4, kStPix = kStTileW * kStTileH;     // reference-pixel tile (UMMA N = 64)
constexpr int kStConsumers = 512;       // warps 0..15: TMEM -> shared copy + look-ups
constexpr int kStPlanWarps = 3;          // warps 17-19: planner (one tile ahead of everybody else)
constexpr int kStThreads = kStConsumers + 32 + 32 * kStPlanWarps;     // warp 16: TMA producer + MMA issuer
constexpr int kStRows = 64;             // band rows per chunk (circular slots: slot = y & 63)
constexpr int kStBox = 32;              // band rows are fetched in runs of 32 pixels (2 KB, four SWIZZLE_64B atoms)
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
  CUtensorMap meas_map[kStMaxMeas][2];  // box {32, 32, 1, 1}: one 32-pixel run of a band row
  const __half* meas_planes[kStMaxMeas][2];   // raw pointers for the direct (fallback) path
  const __half* ref_planes[2];
  const float* pose2[kStMaxMeas];
  const float* pose1;
  const float* K;
  float* out;                           // [B][h][w][D]
  int B, h, w, D, M;
  int tiles_x, tiles_y;
  int qcap;                             // band capacity in pixels (multiple of 32)
  int tmem_cols;                        // TMEM columns claimed: 64 per 128 band pixels of capacity, rounded up to a power of two
  float depth[kStMaxPlanes];            // plane depths, computed on the host in double like the reference (utils.py:59-66)
  long long* timeline;                  // development aid (tools/sweep_timeline.py): clock64 stamps of the first CTAs' phases, or null
};

// same algebra as geometry.cu sweep_matrices (utils.py:51-56): G = K R K^-1, Kt = K t.  The inverses are evaluated in double
// (common.cuh): an fp32 cofactor inverse moves the sampling positions by ~1e-4 px, which the 2e-5 parity bound of the 3-term
// mode does not admit (measured: golden test fails)
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

constexpr int kStMaxChunks = 32;
constexpr int kStMaxRuns = 32;          // 1024 band pixels / 32
struct StChunk {         // one chunk of consecutive planes of one frame
  short m, d0, nd, band, total_q, next_band;       // next_band: index of the next chunk with a band (or -1)
  short row_q[kStRows];            // q index of pixel x on the band row in slot (y & 63) = row_q + x
  short run_x[kStMaxRuns];         // band pixels [32 i, 32 i + 32) are the pixels x = run_x[i] .. + 31 of image row run_y[i]
  short run_y[kStMaxRuns];
};

struct StTileCtx {       // everything the main loop needs to know about one tile; double-buffered, written by the planner warp
  int b, v0, u0, tw, th, n_chunks, first_band, pad;
  float G[kStMaxMeas][12];
  StChunk chunk[kStMaxChunks];
};

struct StSmem {          // fixed-size bookkeeping behind the big arrays
  StTileCtx ctx[2];
  int frame_n[kStMaxMeas], frame_fail[kStMaxMeas], frame_stuck[kStMaxMeas];      // planner scratch
  int plan[kStPlanWarps][2 * kStRows + 4];        // planner scratch per warp: row xmin[64], xmax[64], ylo, yhi, degenerate of the chunk in work
  int any_fail;
  unsigned long long bar_ref, bar_go, bar_mma, bar_band, bar_ctx_ready[2], bar_ctx_free[2];
  uint32_t tmem_slot;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kStConsumers) : "memory"); }
__device__ __forceinline__ void planner_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(32 * kStPlanWarps) : "memory"); }

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
// centre per plane.  Only a first guess -- the planner verifies the real band and shortens the chunks when it does not fit.
__device__ __forceinline__ float st_band_estimate(float n, float dx, float dy) {
  const float rows = (float)(kStTileH + 2) + dy * n;
  const float planes_per_row = fminf(n, (float)(kStTileH + 2) / fmaxf(dy, 1e-6f));
  const float width = (float)(kStTileW + 1) + dx * planes_per_row;
  return rows * (float)kStBox * ceilf(width * (1.f / kStBox));
}

#define ST_STAMP(slot)                                                                                   \
  do {                                                                                                   \
    if (p.timeline && blockIdx.x < 8) p.timeline[blockIdx.x * 64 + (slot)] = clock64();                   \
  } while (0)

// ---- the planner (kStPlanWarps warps, pt = thread index among them): everything about tile `tile` that the main loop needs, into
// ctx / kd.  Runs one tile ahead of the producer and the consumers, so its few thousand dependent instructions (fp64 pose algebra,
// per-plane geometry, chunk plan) hide behind the previous tile's look-ups.
__device__ __forceinline__ void st_plan_tile(const SweepTcParams& p, StSmem* sm, StTileCtx* ctx, float4* kd_out, int4* s_box, int tile, int pt,
                                             float sx, float sy, float wf, float hf) {
  constexpr int kPT = 32 * kStPlanWarps;
  const int lane = pt & 31, pw = pt >> 5;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int b = tile / tiles_per_img;
  const int t_in = tile - b * tiles_per_img;
  const int v0 = (t_in / p.tiles_x) * kStTileH, u0 = (t_in % p.tiles_x) * kStTileW;
  const int tw = min(kStTileW, p.w - u0), th = min(kStTileH, p.h - v0);      // valid extent of this tile
  if (pt < p.M) {
    float G[9], Kt[3];
    st_matrices(p.pose1 + b * 16, p.pose2[pt] + b * 16, p.K + b * 9, G, Kt);
#pragma unroll
    for (int i = 0; i < 9; ++i) ctx->G[pt][i] = G[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) ctx->G[pt][9 + i] = Kt[i];
  }
  planner_barrier();
  // per-plane geometry of all frames: one thread per (frame, plane, tile corner), the four corners combine with shuffles
  const int MD = p.M * p.D;
  for (int i4 = pt; i4 < ((MD * 4 + 31) & ~31); i4 += kPT) {
    const int i = min(i4 >> 2, MD - 1), c = i4 & 3;
    const int m = i / p.D, d = i - m * p.D;
    const float* G = ctx->G[m];
    const float this_depth = p.depth[d];                                                     // utils.py:66
    const float4 kd = make_float4(G[9] / this_depth, G[10] / this_depth, G[11] / this_depth, 0.f);      // utils.py:68
    // bounding box of the tile's image on this plane: a homography with a denominator of one sign maps the (convex) tile
    // onto a convex quadrilateral, so the four corners bound every sample; +-1e-3 px absorbs fp32 rounding
    const float cu = (float)(u0 + ((c & 1) ? tw - 1 : 0)), cv = (float)(v0 + ((c >> 1) ? th - 1 : 0));
    float xs, ys, den;
    st_position(G, kd, cu, cv, sx, sy, wf, hf, xs, ys, den);
    int pos = (den > 1e-6f), neg = (den < -1e-6f);
    float xmn = xs, xmx = xs, ymn = ys, ymx = ys;
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      xmn = fminf(xmn, __shfl_xor_sync(0xffffffffu, xmn, o)); xmx = fmaxf(xmx, __shfl_xor_sync(0xffffffffu, xmx, o));
      ymn = fminf(ymn, __shfl_xor_sync(0xffffffffu, ymn, o)); ymx = fmaxf(ymx, __shfl_xor_sync(0xffffffffu, ymx, o));
      pos += __shfl_xor_sync(0xffffffffu, pos, o); neg += __shfl_xor_sync(0xffffffffu, neg, o);
    }
    if (c == 0 && (i4 >> 2) < MD) {
      int4 bb;
      bb.x = (int)floorf(xmn - 1e-3f); bb.y = (int)floorf(xmx + 1e-3f) + 1;
      bb.z = (int)floorf(ymn - 1e-3f); bb.w = (int)floorf(ymx + 1e-3f) + 1;
      if ((pos != 4 && neg != 4) || bb.w - bb.z >= kStRows) { bb.x = 1; bb.y = 0; }       // degenerate on this tile: direct path
      kd_out[i] = kd;
      s_box[i] = bb;
    }
  }
  planner_barrier();
  // first guess of the planes per chunk of every frame from the motion of the tile centre between the first and the last plane:
  // lane j tries j + 1 uniform chunks, the smallest count whose estimated band fits wins
  const int budget = max(1, kStMaxChunks / p.M);                  // an equal share of the chunk list per frame
  for (int m = pw; m < p.M; m += kStPlanWarps) {
    const float* G = ctx->G[m];
    const float uc = (float)u0 + 0.5f * (float)(tw - 1), vc = (float)v0 + 0.5f * (float)(th - 1);
    float xa, ya, xb, yb, den;
    st_position(G, kd_out[m * p.D], uc, vc, sx, sy, wf, hf, xa, ya, den);
    st_position(G, kd_out[m * p.D + p.D - 1], uc, vc, sx, sy, wf, hf, xb, yb, den);
    const float dx = fabsf(xb - xa) / (float)(p.D - 1), dy = fabsf(yb - ya) / (float)(p.D - 1);
    const int nch_try = lane + 1;
    const int n_try = (p.D + nch_try - 1) / nch_try;
    const bool ok = nch_try >= budget || st_band_estimate((float)n_try, dx, dy) <= (float)p.qcap;
    const unsigned fits_mask = __ballot_sync(0xffffffffu, ok);
    const int nch = fits_mask ? __ffs(fits_mask) : min(budget, 32);
    if (lane == 0) {
      sm->frame_n[m] = (p.D + nch - 1) / nch;
      sm->frame_fail[m] = 0;
      sm->frame_stuck[m] = 0;
    }
  }
  if (pt == 0) sm->any_fail = 0;
  planner_barrier();
  // chunk plan: uniform chunks per frame, every chunk verified against the real per-plane boxes by one warp: its lanes merge the
  // planes' boxes into the chunk's 64 circular row slots (slot = y & 63) with shared-memory atomics, then the rows become 32-pixel
  // runs by prefix sum over the lanes (two slots each).  A frame with a chunk that does not fit is re-planned with shorter chunks
  // as long as the chunk list has room -- what still does not fit then takes the direct path.
  int n_chunks = 0;
  while (true) {              // terminates: a failing frame strictly shortens its chunks until it is stuck, where nothing is flagged any more
    n_chunks = 0;
    for (int m = 0; m < p.M; ++m) {
      const int n = sm->frame_n[m];
      for (int d0 = 0; d0 < p.D; d0 += n, ++n_chunks) {
        if ((n_chunks % kStPlanWarps) != pw) continue;            // chunks are dealt round-robin to the planner warps
        const int nd = min(n, p.D - d0);
        StChunk* ch = &ctx->chunk[n_chunks];
        int* pl_ = sm->plan[pw];
        for (int j = lane; j < 2 * kStRows + 4; j += 32) pl_[j] = (j < kStRows || j == 2 * kStRows) ? INT_MAX : ((j == 2 * kStRows + 2) ? 0 : INT_MIN);
        __syncwarp();
        for (int dd = lane; dd < nd; dd += 32) {
          const int4 bb = s_box[m * p.D + d0 + dd];
          if (bb.x > bb.y) { pl_[2 * kStRows + 2] = 1; continue; }       // degenerate plane in this chunk
          atomicMin(&pl_[2 * kStRows], bb.z);
          atomicMax(&pl_[2 * kStRows + 1], bb.w);
          for (int y = bb.z; y <= bb.w; ++y) {
            atomicMin(&pl_[y & (kStRows - 1)], bb.x);
            atomicMax(&pl_[kStRows + (y & (kStRows - 1))], bb.y);
          }
        }
        __syncwarp();
        const int xmn0 = pl_[2 * lane], xmx0 = pl_[kStRows + 2 * lane], xmn1 = pl_[2 * lane + 1], xmx1 = pl_[kStRows + 2 * lane + 1];
        const int ylo = pl_[2 * kStRows], yhi = pl_[2 * kStRows + 1], degenerate = pl_[2 * kStRows + 2];
        __syncwarp();
        const int n0 = (!degenerate && xmx0 >= xmn0) ? (xmx0 - xmn0 + kStBox) / kStBox : 0;
        const int n1 = (!degenerate && xmx1 >= xmn1) ? (xmx1 - xmn1 + kStBox) / kStBox : 0;
        int incl = n0 + n1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const int total_q = __shfl_sync(0xffffffffu, incl, 31) * kStBox;
        const int r0 = incl - n0 - n1, r1 = r0 + n0;              // first 32-pixel run of the two rows
        const bool fits = !degenerate && yhi - ylo + 1 <= kStRows && total_q <= p.qcap && total_q > 0;
        if (fits) {
          ch->row_q[2 * lane] = (short)(n0 ? r0 * kStBox - xmn0 : 0);
          ch->row_q[2 * lane + 1] = (short)(n1 ? r1 * kStBox - xmn1 : 0);
          const int y0 = ylo + ((2 * lane - ylo) & (kStRows - 1)), y1 = ylo + ((2 * lane + 1 - ylo) & (kStRows - 1));
          for (int j = 0; j < n0; ++j) { ch->run_x[r0 + j] = (short)(xmn0 + kStBox * j); ch->run_y[r0 + j] = (short)y0; }
          for (int j = 0; j < n1; ++j) { ch->run_x[r1 + j] = (short)(xmn1 + kStBox * j); ch->run_y[r1 + j] = (short)y1; }
        }
        if (lane == 0) {
          ch->m = (short)m; ch->d0 = (short)d0; ch->nd = (short)nd;
          ch->band = fits ? 1 : 0;
          ch->total_q = (short)(fits ? total_q : 0);
          if (!fits && !degenerate && !sm->frame_stuck[m] && nd > 1) { sm->frame_fail[m] = 1; sm->any_fail = 1; }      // shorter chunks may fit
        }
      }
    }
    planner_barrier();
    const int any_fail = sm->any_fail;
    planner_barrier();
    if (!any_fail) break;
    if (pt == 0) {
      // shorten the chunks of the failing frames while the chunk list has room (frames share it); a frame that cannot
      // shrink any further is stuck: its chunks that do not fit take the direct path
      for (int m = 0; m < p.M; ++m) {
        if (!sm->frame_fail[m]) continue;
        sm->frame_fail[m] = 0;
        const int n_new = max(1, (sm->frame_n[m] * 3) >> 2);
        int total = 0;
        for (int j = 0; j < p.M; ++j) { const int nj = (j == m) ? n_new : sm->frame_n[j]; total += (p.D + nj - 1) / nj; }
        if (n_new < sm->frame_n[m] && total <= kStMaxChunks) sm->frame_n[m] = n_new;
        else sm->frame_stuck[m] = 1;
      }
      sm->any_fail = 0;
    }
    planner_barrier();
  }
  if (pt == 0) {          // link the band chunks, publish the tile
    int next = -1;
    for (int k = n_chunks - 1; k >= 0; --k) {
      ctx->chunk[k].next_band = (short)next;
      if (ctx->chunk[k].band) next = k;
    }
    ctx->first_band = next;
    ctx->n_chunks = n_chunks;
    ctx->b = b; ctx->v0 = v0; ctx->u0 = u0; ctx->tw = tw; ctx->th = th;
  }
  planner_barrier();
}

// Persistent kernel: CTA c processes tiles c, c + gridDim.x, ...  Warp roles: 0-15 consumers, 16 producer (TMA + MMA), 17-20 planner.
template <int TERMS>
__global__ void __launch_bounds__(kStThreads, 1) plane_sweep_tc_kernel(const __grid_constant__ SweepTcParams p) {
  using Cfg = StCfg<TERMS>;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  // layout: [band hi (+ lo)] [ref tile hi (+ lo)] [S] [acc] [kd x 2] [plane boxes] [bookkeeping]
  const uint32_t band_addr = base;                                               // hi at +0, lo at +qcap*64
  const uint32_t band_bytes = (uint32_t)p.qcap * Cfg::kBandBytesPerQ;            // multiple of 2048
  const uint32_t ref_addr = base + band_bytes;                                   // 64 rows x 64 B per plane
  uint8_t* S = base_ptr + band_bytes + (TERMS == 3 ? 2 : 1) * kStPix * 64;
  float* acc = reinterpret_cast<float*>(S + (size_t)p.qcap * Cfg::kSPitchBytes);
  const int MD = p.M * p.D;
  float4* s_kd = reinterpret_cast<float4*>(acc + ((p.D * kStAccPitch + 3) & ~3));       // [2][M][D]: homography terms of the tile in work / in planning
  int4* s_box = reinterpret_cast<int4*>(s_kd + 2 * MD);                                 // [M][D] {xl, xh, yl, yh}; xl > xh: degenerate (planner only)
  StSmem* sm = reinterpret_cast<StSmem*>(s_box + MD);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool producer = warp == kStConsumers / 32, planner = warp > kStConsumers / 32;
  const int pt = tid - (kStConsumers + 32);          // thread index among the planner warps
  const int tiles_total = p.B * p.tiles_x * p.tiles_y;

  const uint32_t bar_ref = smem_u32(&sm->bar_ref), bar_go = smem_u32(&sm->bar_go), bar_mma = smem_u32(&sm->bar_mma);
  const uint32_t bar_band = smem_u32(&sm->bar_band);
  const uint32_t bar_ready0 = smem_u32(&sm->bar_ctx_ready[0]), bar_free0 = smem_u32(&sm->bar_ctx_free[0]);
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.ref_map[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.meas_map[0][0]) : "memory");
    mbar_init(bar_band, 1);
    mbar_init(bar_ref, 1);
    mbar_init(bar_go, 1);
    mbar_init(bar_mma, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_ready0 + 8u * i, 1); mbar_init(bar_free0 + 8u * i, 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (producer) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_slot)), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();            // barrier objects initialised (thread 0) before anybody arms them; TMEM address published
  tc_fence_after();
  const uint32_t tmem_base = sm->tmem_slot;
  pdl_wait();
  const float sx = (float)(p.w - 1) / (float)p.w, sy = (float)(p.h - 1) / (float)p.h;      // align_corners "shrink" (App. A.1)
  const float wf = (float)p.w, hf = (float)p.h;

  if (planner) {
    // =============================== planner: one tile ahead ===============================
    int i = 0;
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++i) {
      const int s_ = i & 1;
      if (i >= 2) mbar_wait(bar_free0 + 8u * s_, (uint32_t)(((i >> 1) - 1) & 1));      // producer and consumers are done with tile i - 2
      if (pt == 0 && i == 0) ST_STAMP(1);
      st_plan_tile(p, sm, &sm->ctx[s_], s_kd + s_ * MD, s_box, tile, pt, sx, sy, wf, hf);
      if (pt == 0 && i == 0) ST_STAMP(5);
      if (pt == 0) mbar_arrive(bar_ready0 + 8u * s_);           // release: the tile context is complete
    }
  } else if (producer) {
    // =============================== TMA producer + MMA issuer: S = band . tile^T, one band chunk ahead of the look-ups ===============================
    const uint32_t hi_word = umma_hi_word(512u, 4u);                   // SBO = 8 rows x 64 B, SWIZZLE_64B
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kStPix >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f32 += f16 x f16, K-major, N=64, M=128
    const uint32_t b_hi = umma_lo_word(ref_addr, 16), b_lo = umma_lo_word(ref_addr + kStPix * 64, 16);
    int n_band = 0, n_go = 0, i = 0;
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++i) {
      const int s_ = i & 1;
      mbar_wait(bar_ready0 + 8u * s_, (uint32_t)((i >> 1) & 1));
      const StTileCtx* ctx = &sm->ctx[s_];
      // reference tile of this tile: the buffer is free once the previous tile's last MMAs have completed
      if (n_band > 0) mbar_wait(bar_mma, (uint32_t)((n_band - 1) & 1));
      if (lane == 0) {
        mbar_expect_tx(bar_ref, (TERMS == 3 ? 2u : 1u) * kStPix * 64u);
        tma_load_4d(ref_addr, &p.ref_map[0], bar_ref, 0, ctx->u0, ctx->v0, ctx->b);
        if (TERMS == 3) tma_load_4d(ref_addr + kStPix * 64, &p.ref_map[1], bar_ref, 0, ctx->u0, ctx->v0, ctx->b);
      }
      const int b = ctx->b;
      if (ctx->first_band < 0) {
        // no band chunk in this tile: still consume the consumers' "reference tile seen" arrival to stay in lock step
        mbar_wait(bar_go, (uint32_t)(n_go & 1));
        ++n_go;
      }
      for (int k = ctx->first_band; k >= 0; k = ctx->chunk[k].next_band) {
        const StChunk* ch = &ctx->chunk[k];
        const int total_q = ch->total_q, m = ch->m;
        // ---- band rows -> shared memory by TMA, 2 KB runs; pixels outside the image are zero-filled by the TMA unit.  The band
        // buffer is free as soon as the previous band chunk's MMAs have completed, i.e. before its look-ups even start.
        if (n_band > 0) mbar_wait(bar_mma, (uint32_t)((n_band - 1) & 1));
        if (lane == 0) mbar_expect_tx(bar_band, (uint32_t)total_q * (uint32_t)Cfg::kBandBytesPerQ);
        __syncwarp();
        for (int run = lane; run * kStBox < total_q; run += 32) {
          const uint32_t dst = band_addr + (uint32_t)run * 2048u;
          tma_load_4d(dst, &p.meas_map[m][0], bar_band, 0, ch->run_x[run], ch->run_y[run], b);
          if (TERMS == 3) tma_load_4d(dst + (uint32_t)p.qcap * 64u, &p.meas_map[m][1], bar_band, 0, ch->run_x[run], ch->run_y[run], b);
        }
        if (lane == 0) {
          mbar_wait(bar_band, (uint32_t)(n_band & 1));
          // bar_go: per tile, arrival 0 = reference tile landed (and pre-scaled), arrival j = accumulators of its band chunk j-1 drained
          mbar_wait(bar_go, (uint32_t)(n_go & 1));
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
        ++n_go;
        __syncwarp();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free0 + 8u * s_);          // this tile's context may be overwritten (once the consumers agree)
    }
  } else {
    // =============================== consumers: TMEM -> S[q][p], then one thread per (pixel, plane) ===============================
    const int pl = tid & (kStPix - 1);                 // pixel of this thread; planes d0 + (tid >> 6), + 8, ...
    const int pty = pl >> 4, ptx = pl & 15;
    const int wq = warp & 3;
    int n_band = 0, i = 0;
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++i) {
      const int s_ = i & 1;
      if (tid == 0 && i == 0) ST_STAMP(2);
      mbar_wait(bar_ready0 + 8u * s_, (uint32_t)((i >> 1) & 1));
      if (tid == 0 && i == 0) ST_STAMP(6);
      const StTileCtx* ctx = &sm->ctx[s_];
      const int b = ctx->b, v0 = ctx->v0, u0 = ctx->u0, tw = ctx->tw, th = ctx->th, n_chunks = ctx->n_chunks;
      const bool pix_valid = (ptx < tw) && (pty < th);
      const float uf = (float)(u0 + min(ptx, tw - 1)), vf = (float)(v0 + min(pty, th - 1));
      int cur_m = -1;
      float b0 = 0.f, b1 = 0.f, b2 = 0.f;
      // ---- the reference tile arrives by TMA.  1-term mode folds the 1/C of the dot-product cost (utils.py:82) into it: 2^-5 is exact
      // in fp16 (features below 2e-3 go subnormal: 1e-7 of their range), so S = band . tile^T leaves the tensor core pre-scaled and
      // fits fp16 with 32x headroom.  The 3-term mode keeps the tile as it is (its lo plane would go subnormal) and scales the sample.
      mbar_wait(bar_ref, (uint32_t)(i & 1));
      if (TERMS == 1 && ctx->first_band >= 0) {
        __half2* rt = reinterpret_cast<__half2*>(base_ptr + band_bytes);
        const __half2 sc = __floats2half2_rn(1.f / 32.f, 1.f / 32.f);
        for (int j = tid; j < kStPix * 16; j += kStConsumers) rt[j] = __hmul2(rt[j], sc);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> visible to the tensor core's reads
      consumer_barrier();
      if (tid == 0) mbar_arrive(bar_go);
      if (tid == 0 && i == 0) ST_STAMP(7);
      for (int k = 0; k < n_chunks; ++k) {
        const StChunk* ch = &ctx->chunk[k];
        const int m = ch->m, d0 = ch->d0, nd = ch->nd, is_band = (ch->band == 1), total_q = ch->total_q;
        const float* G = ctx->G[m];
        if (m != cur_m) {
          cur_m = m;
          b0 = fmaf(G[0], uf, fmaf(G[1], vf, G[2]));
          b1 = fmaf(G[3], uf, fmaf(G[4], vf, G[5]));
          b2 = fmaf(G[6], uf, fmaf(G[7], vf, G[8]));
        }
        const float4* kdm = s_kd + s_ * MD + m * p.D;
        if (is_band) {
          if (tid == 0 && i == 0) ST_STAMP(8 + 6 * min(k, 7) + 0);
          mbar_wait(bar_mma, (uint32_t)(n_band & 1));
          ++n_band;
          tc_fence_after();
          if (tid == 0 && i == 0) ST_STAMP(8 + 6 * min(k, 7) + 1);
          const int nxt = ch->next_band;
          // ---- TMEM lane = band pixel q, column = tile pixel p  ->  row q of S (all look-ups of the previous chunk are done:
          // consumer barrier at the end of the loop body)
          const int n_mt = (total_q + 127) >> 7;
          for (int mt = warp >> 2; mt < n_mt; mt += kStConsumers / 128) {
            const int q = mt * 128 + wq * 32 + lane;
            float vals[64];
            tmem_ld64(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * kStPix), vals);
            if (q < total_q) {
              if (TERMS == 1) {
                uint4* dst = reinterpret_cast<uint4*>(S + (size_t)q * Cfg::kSPitchBytes);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  __half2 h[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(vals[8 * j + 2 * e], vals[8 * j + 2 * e + 1]);
                  dst[j] = *reinterpret_cast<const uint4*>(h);
                }
              } else {
                float4* dst = reinterpret_cast<float4*>(S + (size_t)q * Cfg::kSPitchBytes);
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[j] = make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
              }
            }
          }
          if (tid == 0 && i == 0) ST_STAMP(8 + 6 * min(k, 7) + 3);
          tc_fence_before();
          consumer_barrier();
          if (nxt >= 0 && tid == 0) mbar_arrive(bar_go);      // the next band chunk's MMAs may overwrite the accumulators
          if (tid == 0 && i == 0) ST_STAMP(8 + 6 * min(k, 7) + 4);
          // ---- look-ups: four scalars per sample, no bounds tests (clamped positions, zero-filled band).  Planes in groups of four:
          // all loads of a group first, then the accumulator updates (acc and S share an address space: interleaving the stores
          // would serialise the loads of the following samples behind them)
          if (pix_valid) {
            const short* row_q = ch->row_q;
            const int qmax = total_q - 2;
            constexpr int kStep = kStConsumers / kStPix;
            for (int dg = d0 + (tid >> 6); dg < d0 + nd; dg += 4 * kStep) {
              float val[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int d = min(dg + j * kStep, d0 + nd - 1);
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
                float v = fmaf(s11, fx * fy, fmaf(s10, gx * fy, fmaf(s01, fx * gy, s00 * (gx * gy))));
                if (TERMS == 3) v *= (1.f / 32.f);                            // utils.py:82 (/C); at 1 term the reference tile is pre-scaled
                val[j] = v;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int d = dg + j * kStep;
                if (d < d0 + nd) {
                  float* a = acc + d * kStAccPitch + pl;
                  *a = (m == 0) ? val[j] : *a + val[j];                       // summed over the measurement frames (utils.py:102)
                }
              }
            }
          }
        } else {
          // ---- direct path for these planes (band does not fit, or the homography is degenerate on the tile)
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
        if (tid == 0 && i == 0 && is_band) ST_STAMP(8 + 6 * min(k, 7) + 5);
      }
      if (tid == 0) mbar_arrive(bar_free0 + 8u * s_);        // the chunk list / homography tables of this tile are no longer needed
      // ---- coalesced write-out: one warp per pixel, D consecutive floats of the channel-last cost volume
      const bool pow2 = (p.M & (p.M - 1)) == 0;
      const float m_f = (float)p.M, m_inv = 1.f / (float)p.M;             // x * (1/M) == x / M exactly when M is a power of two
      for (int px = warp; px < th * kStTileW; px += kStConsumers / 32) {
        const int py = px >> 4, pxx = px & 15;
        if (pxx >= tw) continue;
        float* o = p.out + (((size_t)b * p.h + v0 + py) * p.w + u0 + pxx) * p.D;
        const float* a = acc + px;
        for (int d = lane; d < p.D; d += 32) {
          const float v = a[d * kStAccPitch];
          o[d] = pow2 ? v * m_inv : v / m_f;                                 // utils.py:105-106
        }
      }
      consumer_barrier();                                    // acc is free for the next tile
      if (tid == 0 && i == 0) ST_STAMP(62);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (producer) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------- host side
static int feature_map(CUtensorMap* out, const void* ptr, int B, int h, int w, int box_w, int box_h) {
  cuuint64_t dims[4] = {32, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
  cuuint64_t strides[3] = {64, (cuuint64_t)w * 64, (cuuint64_t)h * w * 64};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = cached_tensor_map(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (r != CUDA_SUCCESS) {
    set_error("plane_sweep_tc: cuTensorMapEncodeTiled(B=%d h=%d w=%d box=%dx%d) failed: %d", B, h, w, box_w, box_h, (int)r);
    return DVMVS_EINVAL;
  }
  return DVMVS_OK;
}

}  // namespace dvmvs

using namespace dvmvs;

static long long* g_timeline = nullptr;
// development aid: device buffer of 8 x 64 long long that the next plane_sweep_tc launches fill with clock64 stamps (null: off)
extern "C" int dvmvs_plane_sweep_tc_set_timeline(void* device_buffer) {
  g_timeline = (long long*)device_buffer;
  return DVMVS_OK;
}

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
  p.timeline = g_timeline;
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
  const size_t fixed = 1024 + (size_t)(terms == 3 ? 2 : 1) * kStPix * 64 + (size_t)((D * kStAccPitch + 3) & ~3) * 4 + (size_t)M * D * 48 + sizeof(StSmem) + 64;
  const int per_q = (terms == 3) ? StCfg<3>::kSPitchBytes + StCfg<3>::kBandBytesPerQ : StCfg<1>::kSPitchBytes + StCfg<1>::kBandBytesPerQ;
  int qcap = (int)((227 * 1024 - fixed) / per_q) & ~(kStBox - 1);
  qcap = min(qcap, terms == 3 ? StCfg<3>::kMaxQ : StCfg<1>::kMaxQ);
  if (qcap_env >= kStBox && qcap_env <= qcap) qcap = qcap_env & ~(kStBox - 1);
  DVMVS_REQUIRE(qcap >= 2 * kStBox, "plane_sweep_tc: no shared memory left for the band (D=%d, M=%d)", D, M);
  p.qcap = qcap;
  {
    const int need = ((qcap + 127) / 128) * kStPix;
    int cols = 32;
    while (cols < need) cols <<= 1;
    p.tmem_cols = cols;
    DVMVS_REQUIRE(cols <= (terms == 3 ? StCfg<3>::kTmemCols : StCfg<1>::kTmemCols), "plane_sweep_tc: band capacity %d needs %d TMEM columns", qcap, cols);
  }
  const size_t smem = fixed + (size_t)qcap * per_q;
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    cudaFuncSetAttribute(plane_sweep_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(plane_sweep_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  // persistent CTAs: one per SM (the shared-memory footprint admits no second), each walking tiles c, c + grid, ...
  static int n_sm_cache[64] = {0};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  if (dev_id < 0 || dev_id >= 64) dev_id = 0;
  if (n_sm_cache[dev_id] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev_id) != cudaSuccess || n <= 0) n = 148;
    n_sm_cache[dev_id] = n;
  }
  const int tiles_total = B * p.tiles_x * p.tiles_y;
  const int ctas = tiles_total < n_sm_cache[dev_id] ? tiles_total : n_sm_cache[dev_id];
  if (terms == 3) launch_k(plane_sweep_tc_kernel<3>, dim3(ctas), dim3(kStThreads), smem, (cudaStream_t)stream, p);
  else launch_k(plane_sweep_tc_kernel<1>, dim3(ctas), dim3(kStThreads), smem, (cudaStream_t)stream, p);
  return check_launch("plane_sweep_tc_kernel");
}
