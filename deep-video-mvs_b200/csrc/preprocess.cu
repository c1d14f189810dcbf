// Image pre-processing on the device: the per-image host work of the reference's test drivers
// (dvmvs/dataset_loader.py:260-263 load_image, :322-334 PreprocessImage.apply_rgb, fusionnet/run-testing.py:127
// transpose + float + upload) as ONE kernel -- crop, cv2.INTER_LINEAR resize, BGR->RGB, /scale, (x-mean)/std,
// HWC -> CHW.  SURVEY.md section 8 row f2: with the network below 1 ms per keyframe the cv2 resize and the three fp32
// uploads per frame dominate the real pipeline; here the decoded uint8 frame is uploaded once (1/4 of the bytes) and
// everything else happens in HBM.
//
// Arithmetic follows OpenCV's float path of resize(INTER_LINEAR) (the reference converts to float32 before resizing):
//   fx = (float)((dx + 0.5) * (double)src_w / dst_w - 0.5);  sx = floor(fx);  fx -= sx;
//   sx < 0 -> (sx, fx) = (0, 0);   sx >= src_w - 1 -> (sx, fx) = (src_w - 1, 0)        (same for y, rows clamped)
//   horizontal pass  t = S[sx] * (1 - fx) + S[sx + 1] * fx   then vertical pass  t0 * (1 - fy) + t1 * fy, all fp32.
// This is bit-for-bit OpenCV's native path (cv2.ipp.setUseIPP(False)) up to one fused multiply-add in its SIMD vertical
// pass; OpenCV builds that dispatch float resizes to Intel IPP round the interpolation coefficients differently
// (<= 0.009 on the 0..255 scale for white noise) -- tests/test_gpu_parity.py checks both.
// HBM-bound: reads <= 4 source texels per output texel (mostly L2 hits), writes 12 B per output pixel.
#include "common.cuh"

namespace dvmvs {

struct PrepParams {
  const void* src;     // [in_h][in_w][3], uint8 or float
  float* dst;          // [3][out_h][out_w]
  int in_h, in_w, crop_x, crop_y, src_h, src_w, out_h, out_w;
  int swap_rb, normalize;
  double scale_x, scale_y;
  float scale, mean[3], stdv[3];
};

template <typename T>
__global__ void __launch_bounds__(256) preprocess_rgb_kernel(PrepParams p) {
  pdl_launch_dependents();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.out_h * p.out_w) return;
  const int dy = idx / p.out_w, dx = idx - dy * p.out_w;
  float fx = (float)(((double)dx + 0.5) * p.scale_x - 0.5);
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) { sx = 0; fx = 0.f; }
  if (sx >= p.src_w - 1) { sx = p.src_w - 1; fx = 0.f; }
  const int sx1 = min(sx + 1, p.src_w - 1);
  float fy = (float)(((double)dy + 0.5) * p.scale_y - 0.5);
  int sy = (int)floorf(fy);
  fy -= (float)sy;
  const int sy0 = min(max(sy, 0), p.src_h - 1), sy1 = min(max(sy + 1, 0), p.src_h - 1);
  const float ax0 = 1.f - fx, ax1 = fx, by0 = 1.f - fy, by1 = fy;
  pdl_wait();
  const T* s = (const T*)p.src;
  const size_t r0 = ((size_t)(sy0 + p.crop_y) * p.in_w + p.crop_x) * 3, r1 = ((size_t)(sy1 + p.crop_y) * p.in_w + p.crop_x) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int cs = p.swap_rb ? 2 - c : c;      // source channel of output channel c
    const float t0 = __fadd_rn(__fmul_rn((float)s[r0 + (size_t)sx * 3 + cs], ax0), __fmul_rn((float)s[r0 + (size_t)sx1 * 3 + cs], ax1));
    const float t1 = __fadd_rn(__fmul_rn((float)s[r1 + (size_t)sx * 3 + cs], ax0), __fmul_rn((float)s[r1 + (size_t)sx1 * 3 + cs], ax1));
    float v = __fadd_rn(__fmul_rn(t0, by0), __fmul_rn(t1, by1));
    if (p.normalize) v = __fdiv_rn(__fsub_rn(__fdiv_rn(v, p.scale), p.mean[c]), p.stdv[c]);
    p.dst[((size_t)c * p.out_h + dy) * p.out_w + dx] = v;
  }
}

}  // namespace dvmvs

using namespace dvmvs;

extern "C" int dvmvs_preprocess_rgb(const void* image, int is_u8, int swap_rb, int in_h, int in_w, int crop_x, int crop_y,
                                    float* out, int out_h, int out_w, int normalize, float scale, const float* mean3,
                                    const float* std3, dvmvs_stream_t stream) {
  DVMVS_REQUIRE(image && out && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "preprocess_rgb: bad argument");
  DVMVS_REQUIRE(crop_x >= 0 && crop_y >= 0 && in_w - 2 * crop_x > 0 && in_h - 2 * crop_y > 0, "preprocess_rgb: crop (%d, %d) leaves nothing of %dx%d",
                crop_x, crop_y, in_w, in_h);
  DVMVS_REQUIRE(!normalize || (mean3 && std3 && scale != 0.f), "preprocess_rgb: normalisation needs scale, mean and std");
  PrepParams p;
  p.src = image;
  p.dst = out;
  p.in_h = in_h; p.in_w = in_w; p.crop_x = crop_x; p.crop_y = crop_y;
  p.src_h = in_h - 2 * crop_y; p.src_w = in_w - 2 * crop_x;
  p.out_h = out_h; p.out_w = out_w;
  p.swap_rb = swap_rb ? 1 : 0;
  p.normalize = normalize ? 1 : 0;
  p.scale_x = (double)p.src_w / (double)out_w;
  p.scale_y = (double)p.src_h / (double)out_h;
  p.scale = scale;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = normalize ? mean3[c] : 0.f;
    p.stdv[c] = normalize ? std3[c] : 1.f;
  }
  const unsigned blocks = (unsigned)(((size_t)out_h * out_w + 255) / 256);
  if (is_u8) launch_k(preprocess_rgb_kernel<unsigned char>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p);
  else launch_k(preprocess_rgb_kernel<float>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p);
  return check_launch("preprocess_rgb_kernel");
}
