// Backward / loss kernels of the training step.  sm_100a.
//
// Reference behaviour reproduced (paths relative to the reference root):
//   dvmvs/convlstm.py:43-59      MVSLayernormConvLSTMCell gate arithmetic (the function whose derivative is taken here)
//   dvmvs/losses.py:26-82        update_losses / calculate_loss (multi-scale L1 / Huber / L1-inv / L1-rel on valid pixels)
// The forward kernels these pair with live in conv.cu (lstm_gates_kernel) -- the backward kernel recomputes the forward
// values from the saved pre-activations instead of storing six activation tensors.
#include "common.cuh"

namespace dvmvs {

__device__ __forceinline__ float t_celu1(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float t_celu1_grad(float x) { return x > 0.f ? 1.f : expf(x); }
__device__ __forceinline__ float t_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int kGateWarps = 8;

// =====================================================================================================
// ConvLSTM gate backward.  Block = 32 channels (lanes) x kGateWarps warps striding over the h*w positions, like the forward.
//   i,f,o = sigmoid(a_i,a_f,a_o);  n = LN_hw(a_g);  gg = celu(n);  cp = f*c + i*gg;  cn = LN_hw(cp);  h = o*celu(cn)
// inputs : gates (B,hw,4C) pre-activations in i,f,o,g order, c_in (B,hw,C), grad_h, grad_c (B,hw,C; grad_c may be null)
// outputs: grad_gates (B,hw,4C), grad_c_in (B,hw,C)
// LayerNorm backward over the positions p of one (b, channel):  dx = rstd * (dy - mean(dy) - y * mean(dy * y)).
// =====================================================================================================
template <int PPW>
__global__ void __launch_bounds__(32 * kGateWarps) lstm_gates_backward_kernel(const float* __restrict__ gates, const float* __restrict__ c_in,
                                                                              const float* __restrict__ grad_h, const float* __restrict__ grad_c,
                                                                              float* __restrict__ grad_gates, float* __restrict__ grad_c_in,
                                                                              int hw, int C) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_red[kGateWarps * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int b = blockIdx.y;
  const float inv_n = 1.f / (float)hw;
  auto block_sum = [&](float v) -> float {
    __syncthreads();
    s_red[warp * 32 + lane] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kGateWarps; ++i) t += s_red[i * 32 + lane];
    return t;
  };

  float ai[PPW], af[PPW], ao[PPW], ag[PPW], vc[PPW], gh[PPW], gc[PPW];
  bool ok[PPW];
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
    const int p = warp + j * kGateWarps;
    ok[j] = p < hw;
    const size_t pp = (size_t)b * hw + (ok[j] ? p : 0);
    const float* gp = gates + pp * 4 * C;
    ai[j] = ok[j] ? gp[c] : 0.f;
    af[j] = ok[j] ? gp[C + c] : 0.f;
    ao[j] = ok[j] ? gp[2 * C + c] : 0.f;
    ag[j] = ok[j] ? gp[3 * C + c] : 0.f;
    vc[j] = ok[j] ? c_in[pp * C + c] : 0.f;
    gh[j] = ok[j] ? grad_h[pp * C + c] : 0.f;
    gc[j] = (ok[j] && grad_c != nullptr) ? grad_c[pp * C + c] : 0.f;
  }
  // ---- forward recomputation (same two-pass statistics as lstm_gates_kernel)
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PPW; ++j) s += ag[j];
  const float mean_g = block_sum(s) * inv_n;
  s = 0.f;
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
    const float d = ag[j] - mean_g;
    if (ok[j]) s += d * d;
  }
  const float rstd_g = rsqrtf(block_sum(s) * inv_n + 1e-5f);
  float cp[PPW];
  s = 0.f;
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
    ag[j] = (ag[j] - mean_g) * rstd_g;                                   // ag now holds n = LN(a_g)
    ai[j] = t_sigmoid(ai[j]);                                            // gate values
    af[j] = t_sigmoid(af[j]);
    ao[j] = t_sigmoid(ao[j]);
    cp[j] = ok[j] ? af[j] * vc[j] + ai[j] * t_celu1(ag[j]) : 0.f;
    s += cp[j];
  }
  const float mean_c = block_sum(s) * inv_n;
  s = 0.f;
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
    const float d = cp[j] - mean_c;
    if (ok[j]) s += d * d;
  }
  const float rstd_c = rsqrtf(block_sum(s) * inv_n + 1e-5f);
  // ---- backward
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
    const float cn = (cp[j] - mean_c) * rstd_c;
    const float d_o = gh[j] * t_celu1(cn);
    const float dcn = ok[j] ? gh[j] * ao[j] * t_celu1_grad(cn) + gc[j] : 0.f;
    gh[j] = d_o * ao[j] * (1.f - ao[j]);                                  // gh now holds grad a_o
    gc[j] = dcn;                                                          // gc now holds d cn
    cp[j] = cn;                                                           // cp now holds cn
    s1 += dcn;
    s2 += dcn * cn;
  }
  const float m1 = block_sum(s1) * inv_n;
  const float m2 = block_sum(s2) * inv_n;
  s1 = 0.f;
  s2 = 0.f;
  float dn[PPW];
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
    const float dcp = ok[j] ? rstd_c * (gc[j] - m1 - cp[j] * m2) : 0.f;   // d (f*c + i*gg)
    const float gg = t_celu1(ag[j]);
    gc[j] = dcp * af[j];                                                  // gc now holds grad c_in
    cp[j] = dcp * vc[j] * af[j] * (1.f - af[j]);                          // cp now holds grad a_f
    vc[j] = dcp * gg * ai[j] * (1.f - ai[j]);                             // vc now holds grad a_i
    dn[j] = dcp * ai[j] * t_celu1_grad(ag[j]);                            // d n
    s1 += dn[j];
    s2 += dn[j] * ag[j];
  }
  const float k1 = block_sum(s1) * inv_n;
  const float k2 = block_sum(s2) * inv_n;
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
    if (!ok[j]) continue;
    const size_t pp = (size_t)b * hw + warp + j * kGateWarps;
    float* gg_out = grad_gates + pp * 4 * C;
    gg_out[c] = vc[j];
    gg_out[C + c] = cp[j];
    gg_out[2 * C + c] = gh[j];
    gg_out[3 * C + c] = rstd_g * (dn[j] - k1 - ag[j] * k2);
    grad_c_in[pp * C + c] = gc[j];
  }
}

// =====================================================================================================
// Multi-scale depth loss (losses.py:43-82): for every prediction scale j, over pixels whose nearest-down-sampled ground
// truth is non-zero: sum |g-p|, sum smooth_l1(p,g), sum |1/g - 1/p|, sum |g-p|/g, and the valid count.  One launch covers all
// scales (blocks are assigned to scales by a prefix table); sums[j][0..4] accumulate with fp32 atomics (zeroed by the
// entry point).  The backward kernel writes d loss / d p for loss = sum_j weight_j * sums[j][type] / sums[j][4], scaled by
// the upstream gradient read from device memory (no host synchronisation on either side).
// =====================================================================================================
constexpr int kMaxScales = 8;
constexpr int kLossThreads = 256;

struct LossParams {
  const float* pred[kMaxScales];
  float* grad[kMaxScales];
  int hs[kMaxScales], ws[kMaxScales];
  int block0[kMaxScales + 1];      // first block of scale j
  float weight[kMaxScales];
  const float* gt;
  float* sums;                     // [n][5]
  const float* upstream;           // scalar (backward) or null
  int n, B, H, W, type;
};

__device__ __forceinline__ bool loss_locate(const LossParams& q, int& j, int& b, int& y, int& x, size_t& idx) {
  j = 0;
  while (j + 1 < q.n && (int)blockIdx.x >= q.block0[j + 1]) ++j;
  idx = (size_t)(blockIdx.x - q.block0[j]) * kLossThreads + threadIdx.x;
  const size_t per = (size_t)q.hs[j] * q.ws[j];
  if (idx >= per * q.B) return false;
  b = (int)(idx / per);
  const int r = (int)(idx - (size_t)b * per);
  y = r / q.ws[j];
  x = r - y * q.ws[j];
  return true;
}

// nearest-neighbour source index of torch.nn.functional.interpolate(mode='nearest'): min(floor(dst * in/out), in - 1)
__device__ __forceinline__ float loss_groundtruth(const LossParams& q, int j, int b, int y, int x) {
  const float sy = (float)q.H / (float)q.hs[j], sx = (float)q.W / (float)q.ws[j];
  const int yy = min((int)floorf((float)y * sy), q.H - 1), xx = min((int)floorf((float)x * sx), q.W - 1);
  return q.gt[((size_t)b * q.H + yy) * q.W + xx];
}

__global__ void __launch_bounds__(kLossThreads) depth_loss_forward_kernel(LossParams q) {
  pdl_launch_dependents();
  pdl_wait();
  int j, b, y, x;
  size_t idx;
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (loss_locate(q, j, b, y, x, idx)) {
    const float g = loss_groundtruth(q, j, b, y, x);
    if (g != 0.f) {
      const float p = q.pred[j][idx];
      const float d = fabsf(g - p);
      v[0] = d;
      v[1] = d < 1.f ? 0.5f * d * d : d - 0.5f;                          // smooth_l1, beta = 1
      v[2] = fabsf(1.f / g - 1.f / p);
      v[3] = d / g;
      v[4] = 1.f;
    }
  }
  __shared__ float s_red[5][kLossThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float t = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) s_red[k][warp] = t;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kLossThreads / 32; ++i) t += s_red[threadIdx.x][i];
    if (t != 0.f) atomicAdd(q.sums + j * 5 + threadIdx.x, t);
  }
}

__global__ void __launch_bounds__(kLossThreads) depth_loss_backward_kernel(LossParams q) {
  pdl_launch_dependents();
  pdl_wait();
  int j, b, y, x;
  size_t idx;
  if (!loss_locate(q, j, b, y, x, idx)) return;
  const float g = loss_groundtruth(q, j, b, y, x);
  float out = 0.f;
  if (g != 0.f) {
    const float p = q.pred[j][idx];
    const float diff = p - g;
    const float sgn = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
    float dl;
    if (q.type == DVMVS_LOSS_L1) dl = sgn;
    else if (q.type == DVMVS_LOSS_HUBER) dl = fabsf(diff) < 1.f ? diff : sgn;
    else if (q.type == DVMVS_LOSS_L1_INV) {
      const float e = 1.f / g - 1.f / p;                                  // d|e|/dp = sign(e) / p^2
      dl = ((e > 0.f) ? 1.f : ((e < 0.f) ? -1.f : 0.f)) / (p * p);
    } else dl = sgn / g;
    out = q.upstream[0] * q.weight[j] / q.sums[j * 5 + 4] * dl;
  }
  q.grad[j][idx] = out;
}

static int fill_loss_params(LossParams& q, const float* const* preds, float* const* grads, const int* hs, const int* ws, const float* weights,
                            int n, int B, int H, int W) {
  int blocks = 0;
  for (int j = 0; j < n; ++j) {
    q.pred[j] = preds[j];
    q.grad[j] = grads ? grads[j] : nullptr;
    q.hs[j] = hs[j];
    q.ws[j] = ws[j];
    q.weight[j] = weights ? weights[j] : 1.f;
    q.block0[j] = blocks;
    blocks += (int)(((size_t)B * hs[j] * ws[j] + kLossThreads - 1) / kLossThreads);
  }
  q.block0[n] = blocks;
  q.n = n; q.B = B; q.H = H; q.W = W;
  return blocks;
}

}  // namespace dvmvs

using namespace dvmvs;

extern "C" int dvmvs_lstm_gates_backward(const float* gates, const float* c_in, const float* grad_h, const float* grad_c, float* grad_gates,
                                         float* grad_c_in, int B, int h, int w, int C, dvmvs_stream_t stream) {
  DVMVS_REQUIRE(gates && c_in && grad_h && grad_gates && grad_c_in, "lstm_gates_backward: null pointer");
  DVMVS_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && C % 32 == 0, "lstm_gates_backward: bad shape (C must be a multiple of 32)");
  const int hw = h * w;
  const int ppw = (hw + kGateWarps - 1) / kGateWarps;
  DVMVS_REQUIRE(ppw <= 16, "lstm_gates_backward: h*w=%d too large (training bottleneck maps have up to 128 positions)", hw);
  dim3 grid(C / 32, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (ppw <= 2) launch_k(lstm_gates_backward_kernel<2>, grid, dim3(32 * kGateWarps), 0, st, gates, c_in, grad_h, grad_c, grad_gates, grad_c_in, hw, C);
  else if (ppw <= 8) launch_k(lstm_gates_backward_kernel<8>, grid, dim3(32 * kGateWarps), 0, st, gates, c_in, grad_h, grad_c, grad_gates, grad_c_in, hw, C);
  else launch_k(lstm_gates_backward_kernel<16>, grid, dim3(32 * kGateWarps), 0, st, gates, c_in, grad_h, grad_c, grad_gates, grad_c_in, hw, C);
  return check_launch("lstm_gates_backward_kernel");
}

extern "C" int dvmvs_depth_loss_forward(const float* const* preds_host, const int* hs_host, const int* ws_host, int n_scales,
                                        const float* groundtruth, float* sums, int B, int H, int W, dvmvs_stream_t stream) {
  DVMVS_REQUIRE(preds_host && hs_host && ws_host && groundtruth && sums, "depth_loss_forward: null pointer");
  DVMVS_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "depth_loss_forward: n_scales=%d outside [1,%d]", n_scales, kMaxScales);
  DVMVS_REQUIRE(B > 0 && H > 0 && W > 0, "depth_loss_forward: bad shape");
  for (int j = 0; j < n_scales; ++j) DVMVS_REQUIRE(preds_host[j] && hs_host[j] > 0 && ws_host[j] > 0, "depth_loss_forward: bad scale %d", j);
  LossParams q = {};
  const int blocks = fill_loss_params(q, preds_host, nullptr, hs_host, ws_host, nullptr, n_scales, B, H, W);
  q.gt = groundtruth;
  q.sums = sums;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(sums, 0, (size_t)n_scales * 5 * sizeof(float), st) != cudaSuccess) {
    set_error("depth_loss_forward: memset failed");
    return DVMVS_ELAUNCH;
  }
  launch_k(depth_loss_forward_kernel, dim3(blocks), dim3(kLossThreads), 0, st, q);
  return check_launch("depth_loss_forward_kernel");
}

extern "C" int dvmvs_depth_loss_backward(const float* const* preds_host, float* const* grads_host, const int* hs_host, const int* ws_host,
                                         const float* weights_host, int n_scales, const float* groundtruth, const float* sums,
                                         const float* upstream, int loss_type, int B, int H, int W, dvmvs_stream_t stream) {
  DVMVS_REQUIRE(preds_host && grads_host && hs_host && ws_host && weights_host && groundtruth && sums && upstream, "depth_loss_backward: null pointer");
  DVMVS_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "depth_loss_backward: n_scales=%d outside [1,%d]", n_scales, kMaxScales);
  DVMVS_REQUIRE(loss_type >= DVMVS_LOSS_L1 && loss_type <= DVMVS_LOSS_HUBER, "depth_loss_backward: bad loss type %d", loss_type);
  for (int j = 0; j < n_scales; ++j) DVMVS_REQUIRE(preds_host[j] && grads_host[j] && hs_host[j] > 0 && ws_host[j] > 0, "depth_loss_backward: bad scale %d", j);
  LossParams q = {};
  const int blocks = fill_loss_params(q, preds_host, grads_host, hs_host, ws_host, weights_host, n_scales, B, H, W);
  q.gt = groundtruth;
  q.sums = const_cast<float*>(sums);
  q.upstream = upstream;
  q.type = loss_type;
  launch_k(depth_loss_backward_kernel, dim3(blocks), dim3(kLossThreads), 0, (cudaStream_t)stream, q);
  return check_launch("depth_loss_backward_kernel");
}
