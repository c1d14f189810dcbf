// Stride-1 k x k convolution on tcgen05 WITHOUT im2col amplification: "halo" implicit GEMM.  sm_100a.
//
// conv_tc_kernel (conv_tc.cu) re-loads the activation tile once per filter tap (25x for 5x5), one TMA row per pixel;
// on the big 32-channel layers (refine, decoder block 4, aggregator0, FPN outputs) it is bound by the TMA request rate.
// Here the activations are stored in a channel-BLOCKED fp16 layout  [plane][B][C/8][H][W][8]  ("NC8HW8"), so that
//   * one TMA box {(8+k-1) pixels x 8 ch = one contiguous row, 16+k-1 rows, KC/8 channel blocks} brings the whole halo
//     of an 8-wide x 16-high output tile into shared memory ONCE per KC-channel group, in rows of (8+k-1)*16 bytes;
//   * in shared memory [channel block][halo y][halo x][8 ch] IS the canonical no-swizzle K-major UMMA layout: a core
//     matrix = 8 x-consecutive pixels x 16 bytes, the next 8 GEMM rows (= next tile row) lie one halo row further
//     (SBO), the next 8 channels one channel block further (LBO);
//   * a filter tap (ky,kx) is just a different START ADDRESS of the same halo tile: (ky*halo_w + kx)*16 bytes.
// So per KC channels the tile issues k*k*(KC/16) MMAs per term from one resident halo, and only the (tiny) per-tap
// weight blocks stream through a ring (one bulk copy per filter row: all kx taps are contiguous in the packed weights).
// Weights are pre-packed in exactly their shared-memory image [n-tile][k-group][ky][kx][KC/8][BLOCK_N][8].
// The epilogue writes fp32 channel-last and/or the blocked fp16 pair planes (16 bytes per pixel and channel block,
// coalesced over the 8 pixels of a tile row).
//
// Warp roles as in conv_tc_kernel: warp 0 = TMA / bulk-copy producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = epilogue.  fp16 (hi, lo) pairs, 3 terms (hi*hi + lo*hi + hi*lo) or 1 term.
#include <string.h>

#include "tc_ptx.cuh"

namespace dvmvs {

constexpr int kHaloThreads = 192;
constexpr int kHaloTileW = 8, kHaloTileH = 16;
constexpr int kHaloWStages = 4;

struct HaloParams {
  CUtensorMap a_map[3][2];     // [source][hi/lo]: 4-D {W*8, H, C8, B} over the blocked planes
  const __half* w_hi;          // packed weights, see header comment
  const __half* w_lo;
  const __half* w_cat;         // TERMS == 2: hi and lo interleaved per 8-channel block, [..][kc/8][2][block_n][8]
  int src_groups[3];           // KC-channel groups per source
  int n_src, terms, ksize, pad, kc;           // kc: channels per group (16 or 32)
  int B, Hout, Wout, Cout, c8_out, tiles_x, tiles_y, n_groups;
  const float* bias;
  const float* residual;       // fp32 channel-last, same size (DVMVS_RES_SAME) or null
  float* out_f32;              // [B][H][W][Cout] or null
  __half* out_blk;             // [2][B][Cout/8][H][W][8] or null
  __half* out_nhwc;            // [2][B][H][W][Cout] or null
  int hi_only;                 // fp16 outputs: hi plane only
  int act;
  uint32_t a_bytes, w_bytes;   // per plane: halo tile bytes of one group, weight bytes of one (group, ky) stage
  unsigned long long* dbg;     // optional timeline dump: [cta][8] globaltimer ns at phase boundaries (profiling aid)
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define HALO_MARK(i) do { if (p.dbg) p.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = gtimer(); } while (0)

template <int BLOCK_N, int KSIZE, int KC, int TERMS>
__global__ void __launch_bounds__(kHaloThreads) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  pdl_launch_dependents();
  if (threadIdx.x == 0) HALO_MARK(0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 127u) & ~127u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const int planes = p.terms > 1 ? 2 : 1;
  // layout: A[2 buffers][planes][a_bytes] | W[kHaloWStages][planes][w_bytes] | barriers
  const uint32_t a_buf_bytes = planes * p.a_bytes, w_stage_bytes = planes * p.w_bytes;
  const uint32_t a_base = base, w_base = base + 2 * a_buf_bytes;
  const uint32_t bars = w_base + kHaloWStages * w_stage_bytes;
  auto a_full = [&](int i) { return bars + 8u * i; };
  auto a_empty = [&](int i) { return bars + 8u * (2 + i); };
  auto w_full = [&](int i) { return bars + 8u * (4 + i); };
  auto w_empty = [&](int i) { return bars + 8u * (4 + kHaloWStages + i); };
  const uint32_t tmem_full_bar = bars + 8u * (4 + 2 * kHaloWStages);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + (bars - base) + 8 * (5 + 2 * kHaloWStages));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int b = blockIdx.x / tiles_per_img;
  const int t_in = blockIdx.x - b * tiles_per_img;
  const int oy0 = (t_in / p.tiles_x) * kHaloTileH, ox0 = (t_in % p.tiles_x) * kHaloTileW;
  const int nt = blockIdx.y, n0 = nt * BLOCK_N;
  const int halo_w = kHaloTileW + p.ksize - 1, halo_h = kHaloTileH + p.ksize - 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_src; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.a_map[s][0]) : "memory");
      if (p.terms > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.a_map[s][1]) : "memory");
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full(i), 1);
      mbar_init(a_empty(i), 1);
    }
    for (int i = 0; i < kHaloWStages; ++i) {
      mbar_init(w_full(i), 1);
      mbar_init(w_empty(i), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // TERMS == 2 ("concatenated" three-term product): the weight operand of a K step is the 2*BLOCK_N-row matrix
  // [W_hi ; W_lo], so ONE MMA with the hi activations yields x_hi*w_hi (columns [0, BLOCK_N)) and x_hi*w_lo (columns
  // [BLOCK_N, 2*BLOCK_N)); a second, BLOCK_N-wide MMA adds x_lo*w_hi to the first half; the epilogue sums the halves.
  // Same three products as TERMS == 3 with two instead of three passes over the activation tile (the shared-memory read
  // of the 128-row A operand is what bounds these small-N tiles) and two instead of three MMA issues.
  constexpr bool CAT = (TERMS == 2);
  constexpr int kTmemCols = CAT ? 2 * BLOCK_N : (BLOCK_N < 32 ? 32 : BLOCK_N);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) HALO_MARK(1);
  pdl_wait();
  if (threadIdx.x == 0) HALO_MARK(2);

  if (warp == 0) {
    // ============================== producer ==============================
    if (lane == 0) {
      int g = 0, wst = 0;
      uint32_t wphase = 0;
      for (int s = 0; s < p.n_src; ++s) {
        for (int cg = 0; cg < p.src_groups[s]; ++cg, ++g) {
          const int ab = g & 1;
          mbar_wait(a_empty(ab), ((g >> 1) & 1) ^ 1u);
          mbar_expect_tx(a_full(ab), planes * p.a_bytes);
          const uint32_t adst = a_base + ab * a_buf_bytes;
          // box {halo_w*8 elements, halo_h rows, kc/8 channel blocks, 1}; out-of-image rows / columns are zero-filled
          tma_load_4d(adst, &p.a_map[s][0], a_full(ab), (ox0 - p.pad) * 8, oy0 - p.pad, cg * (p.kc / 8), b);
          if (p.terms > 1) tma_load_4d(adst + p.a_bytes, &p.a_map[s][1], a_full(ab), (ox0 - p.pad) * 8, oy0 - p.pad, cg * (p.kc / 8), b);
          for (int ky = 0; ky < p.ksize; ++ky) {
            mbar_wait(w_empty(wst), wphase ^ 1u);
            mbar_expect_tx(w_full(wst), planes * p.w_bytes);
            const size_t woff = ((((size_t)nt * p.n_groups + g) * p.ksize + ky) * (size_t)p.w_bytes) / sizeof(__half);
            const uint32_t wdst = w_base + wst * w_stage_bytes;
            if (CAT) {
              bulk_load(wdst, p.w_cat + 2 * woff, 2 * p.w_bytes, w_full(wst));
            } else {
              bulk_load(wdst, p.w_hi + woff, p.w_bytes, w_full(wst));
              if (p.terms > 1) bulk_load(wdst + p.w_bytes, p.w_lo + woff, p.w_bytes, w_full(wst));
            }
            if (++wst == kHaloWStages) { wst = 0; wphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // all offsets below are in 16-byte units (the granularity of the descriptor's address field)
      const uint32_t a_chunk16 = (uint32_t)halo_h * halo_w;          // one 8-channel block of the halo tile
      constexpr uint32_t w_rows = CAT ? 2 * BLOCK_N : BLOCK_N;       // rows of the weight operand of one 8-channel block
      constexpr uint32_t w_chunk16 = w_rows;                         // one 8-channel block of a weight tap: [w_rows][8]
      constexpr uint32_t w_tap16 = (KC / 8) * w_rows;                // one tap: [KC/8][w_rows][8]
      const uint32_t idesc_cat = (1u << 4) | ((uint32_t)((2 * BLOCK_N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_hi_word = umma_hi_word((uint32_t)halo_w * 16, 0);     // SBO = next tile row (next 8 GEMM rows)
      const uint32_t w_hi_word = umma_hi_word(128, 0);                        // SBO = next 8 output channels
      const uint32_t a_lbo = ((a_chunk16 & 0x3FFF) << 16), w_lbo = ((w_chunk16 & 0x3FFF) << 16);
      int g = 0, wst = 0;
      uint32_t wphase = 0, accumulate = 0;
      for (int s = 0; s < p.n_src; ++s) {
        for (int cg = 0; cg < p.src_groups[s]; ++cg, ++g) {
          const int ab = g & 1;
          mbar_wait(a_full(ab), (g >> 1) & 1);
          tc_fence_after();
          if (g == 0) HALO_MARK(3);
          const uint32_t a_plane0 = ((a_base + ab * a_buf_bytes) >> 4) | a_lbo, a_plane1 = a_plane0 + (p.a_bytes >> 4);
#pragma unroll 1
          for (int ky = 0; ky < KSIZE; ++ky) {
            mbar_wait(w_full(wst), wphase);
            tc_fence_after();
            const uint32_t w_plane0 = ((w_base + wst * w_stage_bytes) >> 4) | w_lbo, w_plane1 = w_plane0 + (p.w_bytes >> 4);
            const uint32_t a_row = (uint32_t)ky * halo_w;
#pragma unroll
            for (int kx = 0; kx < KSIZE; ++kx) {
              if constexpr (CAT) {
#pragma unroll
                for (int k2 = 0; k2 < KC / 16; ++k2) {
                  const uint32_t a_off = a_row + kx + 2 * k2 * a_chunk16;
                  const uint32_t w_lo = w_plane0 + kx * w_tap16 + 2 * k2 * w_chunk16;
                  tc_mma_f16_words(tmem_base, a_plane0 + a_off, a_hi_word, w_lo, w_hi_word, idesc_cat, accumulate);   // x_hi * [w_hi ; w_lo]
                  tc_mma_f16_words(tmem_base, a_plane1 + a_off, a_hi_word, w_lo, w_hi_word, idesc, 1u);               // x_lo * w_hi
                  accumulate = 1;
                }
              } else {
#pragma unroll
                for (int term = 0; term < TERMS; ++term) {
#pragma unroll
                  for (int k2 = 0; k2 < KC / 16; ++k2) {
                    const uint32_t a_lo = ((term == 1) ? a_plane1 : a_plane0) + a_row + kx + 2 * k2 * a_chunk16;
                    const uint32_t w_lo = ((term == 2) ? w_plane1 : w_plane0) + kx * w_tap16 + 2 * k2 * w_chunk16;
                    tc_mma_f16_words(tmem_base, a_lo, a_hi_word, w_lo, w_hi_word, idesc, accumulate);
                    accumulate = 1;
                  }
                }
              }
            }
            tc_commit(w_empty(wst));
            if (++wst == kHaloWStages) { wst = 0; wphase ^= 1u; }
          }
          tc_commit(a_empty(ab));
        }
      }
      tc_commit(tmem_full_bar);
      HALO_MARK(4);
    }
  } else {
    // ============================== epilogue ==============================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ty = row >> 3, tx = row & 7;
    const int oy = oy0 + ty, ox = ox0 + tx;
    const bool valid = (oy < p.Hout) && (ox < p.Wout);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) HALO_MARK(5);
    const size_t pix = ((size_t)b * p.Hout + oy) * p.Wout + ox;
    const size_t hw = (size_t)p.Hout * p.Wout;
    const size_t plane_elems = (size_t)p.B * p.c8_out * hw * 8;
    const size_t nhwc_plane = (size_t)p.B * hw * p.Cout;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 8) {
      float v[8];
      tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if constexpr (CAT) {
        float u[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BLOCK_N + c0), u);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += u[e];
      }
      const int cbase = n0 + c0;
      if (!valid || cbase >= p.Cout) continue;
      if (p.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (p.residual) {
        const float* rr = p.residual + pix * p.Cout + cbase;
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(rr)), r1 = __ldg(reinterpret_cast<const float4*>(rr + 4));
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = tc_act(v[e], p.act);
      if (p.out_f32) {
        float* o = p.out_f32 + pix * p.Cout + cbase;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (p.out_blk || p.out_nhwc) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          hi[e] = __float2half_rn(v[e]);
          lo[e] = __float2half_rn(v[e] - __half2float(hi[e]));
        }
        if (p.out_blk) {
          __half* o = p.out_blk + (((size_t)b * p.c8_out + (cbase >> 3)) * hw + (size_t)oy * p.Wout + ox) * 8;
          *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hi);
          if (!p.hi_only) *reinterpret_cast<uint4*>(o + plane_elems) = *reinterpret_cast<const uint4*>(lo);
        }
        if (p.out_nhwc) {
          __half* o = p.out_nhwc + pix * p.Cout + cbase;
          *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hi);
          if (!p.hi_only) *reinterpret_cast<uint4*>(o + nhwc_plane) = *reinterpret_cast<const uint4*>(lo);
        }
      }
    }
    if (threadIdx.x == 64) HALO_MARK(6);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
  if (threadIdx.x == 0) HALO_MARK(7);
}

// fp32 channel-last -> channel blocks [c_offset/8, (c_offset + c_cover)/8) of the blocked fp16 pair planes
// [2][B][C8][H'][W'][8] (the C values of x, then zeros); optional x2 bilinear (align_corners) upsampling on the way.
// One thread per (pixel, 8-channel block): 32-byte reads, one 16-byte store per plane, consecutive threads = consecutive
// pixels of one channel block (coalesced 512-byte stores per warp).  c_offset must be a multiple of 8.
__global__ void split_blocked_kernel(const float* __restrict__ x, __half* __restrict__ planes, int B, int H, int W, int C, int C8,
                                     int upsample, int c_offset, int c_cover) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = upsample ? 2 * H : H, Wo = upsample ? 2 * W : W;
  const size_t hw = (size_t)Ho * Wo;
  const size_t total = (size_t)B * C8 * hw * 8;
  const int nblk = (c_cover + 7) >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * nblk * hw) return;
  const size_t p_in = idx % hw;
  const int cb = (int)((idx / hw) % nblk);
  const int b = (int)(idx / (hw * nblk));
  const int ox = (int)(p_in % Wo), oy = (int)(p_in / Wo);
  const int c0 = cb * 8;                       // first source channel of this block
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  const bool vec = ((C & 3) == 0) && (c0 + 8 <= C);
  if (!upsample) {
    const float* src = x + ((size_t)b * hw + p_in) * C + c0;
    if (vec) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(src)), a1 = __ldg(reinterpret_cast<const float4*>(src + 4));
      v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c0 + e < C) v[e] = __ldg(src + e);
    }
  } else {
    const float sh = (Ho > 1) ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
    const float sw = (Wo > 1) ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    const float fy = sh * oy, fx = sw * ox;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1), x1 = x0 + (x0 < W - 1);
    const float ly1 = fy - y0, lx1 = fx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* bp = x + (size_t)b * H * W * C + c0;
    const float* p00 = bp + ((size_t)y0 * W + x0) * C;
    const float* p01 = bp + ((size_t)y0 * W + x1) * C;
    const float* p10 = bp + ((size_t)y1 * W + x0) * C;
    const float* p11 = bp + ((size_t)y1 * W + x1) * C;
    if (vec) {       // 8 x 16-byte loads instead of 32 scalar ones (this kernel is LSU-issue bound, not bandwidth bound)
      float t00[8], t01[8], t10[8], t11[8];
      *reinterpret_cast<float4*>(t00) = __ldg(reinterpret_cast<const float4*>(p00));
      *reinterpret_cast<float4*>(t00 + 4) = __ldg(reinterpret_cast<const float4*>(p00 + 4));
      *reinterpret_cast<float4*>(t01) = __ldg(reinterpret_cast<const float4*>(p01));
      *reinterpret_cast<float4*>(t01 + 4) = __ldg(reinterpret_cast<const float4*>(p01 + 4));
      *reinterpret_cast<float4*>(t10) = __ldg(reinterpret_cast<const float4*>(p10));
      *reinterpret_cast<float4*>(t10 + 4) = __ldg(reinterpret_cast<const float4*>(p10 + 4));
      *reinterpret_cast<float4*>(t11) = __ldg(reinterpret_cast<const float4*>(p11));
      *reinterpret_cast<float4*>(t11 + 4) = __ldg(reinterpret_cast<const float4*>(p11 + 4));
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = ly0 * (lx0 * t00[e] + lx1 * t01[e]) + ly1 * (lx0 * t10[e] + lx1 * t11[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c0 + e < C) v[e] = ly0 * (lx0 * __ldg(p00 + e) + lx1 * __ldg(p01 + e)) + ly1 * (lx0 * __ldg(p10 + e) + lx1 * __ldg(p11 + e));
    }
  }
  __align__(16) __half hi[8];
  __align__(16) __half lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    hi[e] = __float2half_rn(v[e]);
    lo[e] = __float2half_rn(v[e] - __half2float(hi[e]));
  }
  const size_t o = (((size_t)b * C8 + (c_offset >> 3) + cb) * hw + p_in) * 8;
  *reinterpret_cast<uint4*>(planes + o) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(planes + total + o) = *reinterpret_cast<const uint4*>(lo);
}

static int make_halo_map(CUtensorMap* map, const void* ptr, int B, int H, int W, int C8, int halo_w, int halo_h, int kc8) {
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)C8, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)C8 * H * W * 16};
  cuuint32_t box[4] = {(cuuint32_t)halo_w * 8, (cuuint32_t)halo_h, (cuuint32_t)kc8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = cached_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(blocked activation B=%d H=%d W=%d C8=%d halo %dx%d) failed: %d", B, H, W, C8, halo_w, halo_h, (int)r);
    return DVMVS_EINVAL;
  }
  return DVMVS_OK;
}

template <int BLOCK_N, int KSIZE, int KC, int TERMS>
static int launch_halo(const HaloParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BLOCK_N, KSIZE, KC, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("conv_halo smem attribute: %s", cudaGetErrorString(e)); return DVMVS_ELAUNCH; }
  }
  launch_k(conv_halo_kernel<BLOCK_N, KSIZE, KC, TERMS>, grid, dim3(kHaloThreads), smem, s, p);
  return check_launch("conv_halo_kernel");
}

template <int BLOCK_N, int KSIZE>
static int launch_halo_kt(const HaloParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  if (p.terms == 3 && p.w_cat) {      // concatenated-weights form of the three-term product
    return p.kc == 16 ? launch_halo<BLOCK_N, KSIZE, 16, 2>(p, grid, smem, s) : launch_halo<BLOCK_N, KSIZE, 32, 2>(p, grid, smem, s);
  }
  if (p.kc == 16) return p.terms == 3 ? launch_halo<BLOCK_N, KSIZE, 16, 3>(p, grid, smem, s) : launch_halo<BLOCK_N, KSIZE, 16, 1>(p, grid, smem, s);
  return p.terms == 3 ? launch_halo<BLOCK_N, KSIZE, 32, 3>(p, grid, smem, s) : launch_halo<BLOCK_N, KSIZE, 32, 1>(p, grid, smem, s);
}

static void* g_halo_dbg = nullptr;

}  // namespace dvmvs

using namespace dvmvs;

// profiling aid: device buffer of (#CTAs x 8) uint64 receiving %globaltimer at the phase boundaries of conv_halo_kernel
// (0 start, 1 setup done, 2 dependencies resolved, 3 first halo landed, 4 last MMA issued, 5 accumulator ready,
// 6 epilogue stores issued, 7 exit); NULL switches it off.
extern "C" int dvmvs_debug_set_halo_timeline(void* device_buffer) {
  g_halo_dbg = device_buffer;
  return DVMVS_OK;
}

extern "C" int dvmvs_conv2d_halo(const dvmvs_conv_halo_desc* d, dvmvs_stream_t stream) {
  DVMVS_REQUIRE(d != nullptr, "conv2d_halo: null descriptor");
  DVMVS_REQUIRE(tensor_map_encoder() != nullptr, "conv2d_halo: cuTensorMapEncodeTiled entry point not available");
  DVMVS_REQUIRE(d->n_src >= 1 && d->n_src <= 3, "conv2d_halo: n_src=%d", d->n_src);
  DVMVS_REQUIRE(d->ksize == 3 || d->ksize == 5, "conv2d_halo: ksize=%d (3 or 5)", d->ksize);
  DVMVS_REQUIRE(d->terms == 1 || d->terms == 3, "conv2d_halo: terms=%d", d->terms);
  DVMVS_REQUIRE(d->kc == 16 || d->kc == 32, "conv2d_halo: kc=%d", d->kc);
  DVMVS_REQUIRE(d->block_n == 32 || d->block_n == 64, "conv2d_halo: block_n=%d", d->block_n);
  DVMVS_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cout > 0 && d->Cout % 8 == 0 && d->w_hi && (d->terms == 1 || d->w_lo),
                "conv2d_halo: bad shape / null weights (Cout must be a multiple of 8)");
  DVMVS_REQUIRE(d->out_f32 || d->out_blk || d->out_nhwc, "conv2d_halo: no output");
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.ksize = d->ksize; p.pad = (d->ksize - 1) / 2; p.terms = d->terms; p.kc = d->kc; p.n_src = d->n_src;
  p.B = d->B; p.Hout = d->H; p.Wout = d->W; p.Cout = d->Cout; p.c8_out = d->Cout / 8;
  p.tiles_x = (d->W + kHaloTileW - 1) / kHaloTileW;
  p.tiles_y = (d->H + kHaloTileH - 1) / kHaloTileH;
  const int halo_w = kHaloTileW + d->ksize - 1, halo_h = kHaloTileH + d->ksize - 1;
  int n_groups = 0;
  for (int s = 0; s < d->n_src; ++s) {
    const int C8 = d->src_c8[s];
    DVMVS_REQUIRE(d->src_blk[s] && C8 > 0, "conv2d_halo: source %d null / empty", s);
    DVMVS_REQUIRE((uintptr_t)d->src_blk[s] % 16 == 0, "conv2d_halo: source %d not 16-byte aligned", s);
    p.src_groups[s] = (C8 * 8 + d->kc - 1) / d->kc;
    n_groups += p.src_groups[s];
    const size_t plane = (size_t)d->B * C8 * d->H * d->W * 8;
    int rc = make_halo_map(&p.a_map[s][0], d->src_blk[s], d->B, d->H, d->W, C8, halo_w, halo_h, d->kc / 8);
    if (rc != DVMVS_OK) return rc;
    if (d->terms > 1) {
      rc = make_halo_map(&p.a_map[s][1], (const __half*)d->src_blk[s] + plane, d->B, d->H, d->W, C8, halo_w, halo_h, d->kc / 8);
      if (rc != DVMVS_OK) return rc;
    }
  }
  p.n_groups = n_groups;
  DVMVS_REQUIRE(d->n_groups == n_groups, "conv2d_halo: weights packed for %d channel groups, sources give %d", d->n_groups, n_groups);
  p.a_bytes = (uint32_t)(d->kc / 8) * halo_h * halo_w * 16;
  p.w_bytes = (uint32_t)d->ksize * (d->kc / 8) * d->block_n * 16;
  p.w_hi = (const __half*)d->w_hi; p.w_lo = (const __half*)d->w_lo;
  p.w_cat = (d->terms == 3) ? (const __half*)d->w_cat : nullptr;
  DVMVS_REQUIRE(!p.w_cat || (uintptr_t)p.w_cat % 16 == 0, "conv2d_halo: w_cat not 16-byte aligned");
  p.bias = d->bias; p.residual = d->residual; p.act = d->act;
  p.out_f32 = d->out_f32; p.out_blk = (__half*)d->out_blk; p.out_nhwc = (__half*)d->out_nhwc;
  p.hi_only = d->out_hi_only ? 1 : 0;
  p.dbg = (unsigned long long*)g_halo_dbg;
  const int planes = d->terms > 1 ? 2 : 1;
  const size_t smem = 2 * (size_t)planes * p.a_bytes + kHaloWStages * (size_t)planes * p.w_bytes + 256 + 128;
  DVMVS_REQUIRE(smem <= 227 * 1024, "conv2d_halo: shared memory %zu too large", smem);
  const int n_tiles = (d->Cout + d->block_n - 1) / d->block_n;
  dim3 grid(p.tiles_x * p.tiles_y * d->B, n_tiles, 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->block_n == 32) return d->ksize == 3 ? launch_halo_kt<32, 3>(p, grid, smem, st) : launch_halo_kt<32, 5>(p, grid, smem, st);
  return d->ksize == 3 ? launch_halo_kt<64, 3>(p, grid, smem, st) : launch_halo_kt<64, 5>(p, grid, smem, st);
}

extern "C" int dvmvs_split_blocked(const float* x, void* planes, int B, int H, int W, int C, int C8, int upsample2x, int c_offset,
                                   int c_cover, dvmvs_stream_t stream) {
  DVMVS_REQUIRE(x && planes && B > 0 && H > 0 && W > 0 && C > 0 && C8 > 0, "split_blocked: bad argument");
  DVMVS_REQUIRE(c_offset >= 0 && c_cover >= C && c_offset + c_cover <= C8 * 8, "split_blocked: channel window [%d,+%d) outside %d",
                c_offset, c_cover, C8 * 8);
  DVMVS_REQUIRE(c_offset % 8 == 0, "split_blocked: c_offset must be a multiple of 8 (got %d)", c_offset);
  const size_t total = (size_t)B * H * W * ((c_cover + 7) / 8) * (upsample2x ? 4 : 1);
  launch_k(split_blocked_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, (__half*)planes, B, H, W, C,
           C8, upsample2x, c_offset, c_cover);
  return check_launch("split_blocked_kernel");
}
