// CLIP scorer kernels (reference sd/scorers.py:149-213 -> transformers CLIPImageProcessor + CLIPModel).
//
// Preprocessing is the processor's PIL path restated in integer arithmetic, bit for bit: Pillow's two-pass 8-bit bicubic
// resampler (Resample.c: coefficients scaled by 2^22, int accumulator started at 2^21, >> 22, clip to 0..255, a uint8
// intermediate between the horizontal and the vertical pass), the centre crop, and rescale + normalise as a 3 x 256 table
// computed by the host in the processor's own arithmetic.  The vertical pass writes straight into the patch matrix of the
// patch-embedding GEMM (row = token, column = c * P * P + i * P + j, zero padded to a multiple of 64), so the fp32
// pixel_values tensor never exists.  HBM-bound byte work: coalesced byte reads, 16-bit coalesced writes.
#pragma once
#include "common.cuh"

namespace b200 {

struct ClipPreArgs {
  const uint8_t* img;        // [B, 3, H, W]
  uint8_t* tmp;              // [B, 3, H, S]   horizontal pass, the cropped columns only
  act_t* patches;            // [B * Lp, Kp]   row b * Lp + 1 + py * G + px
  const int32_t* hb;         // [S, 2] (first input column, taps) of output column left + x
  const int32_t* hk;         // [S, hks]
  const int32_t* vb;         // [S, 2] rows
  const int32_t* vk;         // [S, vks]
  const float* lut;          // [3, 256]
  int batch, H, W, S, P, G, Lp, Kp, hks, vks;
};

constexpr int kClipPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8_resample(int ss) {
  ss >>= kClipPrecisionBits;
  return static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
}

// one thread per (b, c, y, x) of the horizontally resampled, column-cropped image
__global__ void __launch_bounds__(256) clip_resize_h_kernel(ClipPreArgs a) {
  const int64_t total = static_cast<int64_t>(a.batch) * 3 * a.H * a.S;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int x = static_cast<int>(i % a.S);
    const int64_t row = i / a.S;                                    // (b * 3 + c) * H + y
    const int xmin = __ldg(a.hb + 2 * x), n = __ldg(a.hb + 2 * x + 1);
    const uint8_t* src = a.img + row * a.W + xmin;
    const int32_t* k = a.hk + static_cast<int64_t>(x) * a.hks;
    int ss = 1 << (kClipPrecisionBits - 1);
    for (int t = 0; t < n; ++t) ss += static_cast<int>(__ldg(src + t)) * __ldg(k + t);
    a.tmp[i] = clip8_resample(ss);
  }
}

// one thread per element of the patch matrix (token row, k): vertical resample of the 14 x 14 x 3 pixels of the patch
__global__ void __launch_bounds__(256) clip_patches_kernel(ClipPreArgs a) {
  const int GG = a.G * a.G, PP = a.P * a.P;
  const int64_t total = static_cast<int64_t>(a.batch) * GG * a.Kp;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int k = static_cast<int>(i % a.Kp);
    const int64_t tok = i / a.Kp;
    const int b = static_cast<int>(tok / GG), g = static_cast<int>(tok % GG);
    float v = 0.f;
    if (k < 3 * PP) {
      const int c = k / PP, r = k % PP;
      const int y = (g / a.G) * a.P + r / a.P, x = (g % a.G) * a.P + r % a.P;
      const int ymin = __ldg(a.vb + 2 * y), n = __ldg(a.vb + 2 * y + 1);
      const uint8_t* src = a.tmp + ((static_cast<int64_t>(b) * 3 + c) * a.H + ymin) * a.S + x;
      const int32_t* kk = a.vk + static_cast<int64_t>(y) * a.vks;
      int ss = 1 << (kClipPrecisionBits - 1);
      for (int t = 0; t < n; ++t) ss += static_cast<int>(__ldg(src + static_cast<int64_t>(t) * a.S)) * __ldg(kk + t);
      v = __ldg(a.lut + c * 256 + clip8_resample(ss));
    }
    a.patches[(static_cast<int64_t>(b) * a.Lp + 1 + g) * a.Kp + k] = f2act(v);
  }
}

// class-token pooling + post-LayerNorm in fp32: out[b, :] = LN(x[b * row_stride + 0 .. C)) -- one CTA per image
__global__ void __launch_bounds__(256) clip_pool_ln_kernel(const act_t* __restrict__ x, int64_t row_stride, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ out, int C, float eps) {
  __shared__ float red[2][8];
  const act_t* xr = x + blockIdx.x * row_stride;
  float s = 0.f, q = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float v = act2f(xr[c]);
    s += v;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < 8; ++w) mean += red[0][w];
  mean /= static_cast<float>(C);
  for (int c = threadIdx.x; c < C; c += 256) {
    const float d = act2f(xr[c]) - mean;
    q += d * d;
  }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = q;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < 8; ++w) var += red[1][w];
  const float rstd = rsqrtf(var / static_cast<float>(C) + eps);
  for (int c = threadIdx.x; c < C; c += 256)
    out[static_cast<int64_t>(blockIdx.x) * C + c] = (act2f(xr[c]) - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
}

// score[b] = sum_d (img[b,d] / |img[b]|) * (txt[b or 0, d] / |txt|)      (sd/scorers.py:178-213) -- one warp per image
__global__ void __launch_bounds__(128) clip_cosine_kernel(const float* __restrict__ img, const float* __restrict__ txt, int txt_rows,
                                                          float* __restrict__ score, int B, int D) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* ir = img + static_cast<int64_t>(b) * D;
  const float* tr = txt + static_cast<int64_t>(txt_rows > 1 ? b : 0) * D;
  float si = 0.f, st = 0.f;
  for (int d = lane; d < D; d += 32) {
    si += ir[d] * ir[d];
    st += tr[d] * tr[d];
  }
  for (int o = 16; o; o >>= 1) {
    si += __shfl_xor_sync(0xffffffffu, si, o);
    st += __shfl_xor_sync(0xffffffffu, st, o);
  }
  const float ni = sqrtf(si), nt = sqrtf(st);
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) acc += (ir[d] / ni) * (tr[d] / nt);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) score[b] = acc;
}

}  // namespace b200
