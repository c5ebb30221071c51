// Leaves of the SD VAE decoder (AutoencoderKL.decode, sd/diffusers/src/diffusers/models/autoencoders/autoencoder_kl.py:287-320,
// vae.py Decoder) that the U-Net kernels do not already cover -- SURVEY.md 8 f1 (VAE decode inside the search loop):
//   softmax_rows_kernel   P = softmax(scale * S) over the key axis, fp32 logits -> bf16 probabilities.  The decoder's single
//                         mid-block attention has ONE head of dimension 512 over 4096 tokens: its accumulator alone (128 x 512
//                         fp32) fills TMEM, so it runs unfused as  S = Q K^T (tcgen05 GEMM, fp32 out) -> this kernel ->
//                         O = P V (tcgen05 GEMM); 17 + 17 GFLOP per image next to the decoder's 2.48 TFLOP.
//   post_quant_kernel     the 1x1 `post_quant_conv` (4 -> 4 channels) on the fp32 NCHW latents
//   image_sums_kernel     decoded image fp32 NHWC [B,H,W,3] -> uint8 quantisation (x*127.5+128, clip, truncate;
//                         pipeline_stable_diffusion.py:1115) -> integer channel sums (BrightnessScorer, sd/scorers.py:25-76)
//                         and, optionally, the uint8 NCHW image for generic scorers.
#pragma once
#include "common.cuh"

namespace b200 {

// one CTA (256 threads) per row; L multiple of 8, L <= 256 * 32
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, act_t* __restrict__ P, int L,
                                                           float scale_log2e) {
  __shared__ float s_red[8];
  __shared__ float s_bcast;
  const int64_t row = blockIdx.x;
  const float* s = S + row * L;
  float v[32];
  float mx = -INFINITY;
  int n = 0;
  for (int c = threadIdx.x * 4; c < L; c += 1024, ++n) {
    const float4 t = *reinterpret_cast<const float4*>(s + c);
    v[4 * n + 0] = t.x; v[4 * n + 1] = t.y; v[4 * n + 2] = t.z; v[4 * n + 3] = t.w;
    mx = fmaxf(fmaxf(mx, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
    s_bcast = m;
  }
  __syncthreads();
  const float mc = s_bcast * scale_log2e;
  float sum = 0.f;
  for (int i = 0; i < 4 * n; ++i) {
    v[i] = ex2_approx(fmaf(v[i], scale_log2e, -mc));
    sum += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];          // fixed order
    s_bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = s_bcast;
  act_t* p = P + row * L;
  n = 0;
  for (int c = threadIdx.x * 4; c < L; c += 1024, ++n) {
    uint2 u;
    u.x = pack_act(v[4 * n + 0] * inv, v[4 * n + 1] * inv);
    u.y = pack_act(v[4 * n + 2] * inv, v[4 * n + 3] * inv);
    *reinterpret_cast<uint2*>(p + c) = u;
  }
}

// out[b, o, p] = bias[o] + sum_i w[o, i] * x[b, i, p]   (C <= 8 channels, fp32 NCHW)
__global__ void __launch_bounds__(256) post_quant_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int B, int C,
                                                         int HW) {
  const int64_t total = static_cast<int64_t>(B) * HW;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / HW, p = i - b * HW;
    float in[8];
    for (int c = 0; c < C; ++c) in[c] = x[(b * C + c) * HW + p];
    for (int o = 0; o < C; ++o) {
      float acc = 0.f;                                   // F.conv2d accumulates the products, then adds the bias
      for (int c = 0; c < C; ++c) acc = fmaf(__ldg(w + o * C + c), in[c], acc);
      out[(b * C + o) * HW + p] = acc + __ldg(bias + o);
    }
  }
}

// grid (chunks, B): img fp32 NHWC [B, HW, C<=4] -> chan_sums[B,4] (+= ; zeroed by the caller) and optional u8 NCHW
__global__ void __launch_bounds__(256) image_sums_kernel(const float* __restrict__ img, uint32_t* __restrict__ chan_sums,
                                                         uint8_t* __restrict__ u8, int C, int HW) {
  const int64_t b = blockIdx.y;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * per, p_end = min(HW, p_begin + per);
  uint32_t local[4] = {0, 0, 0, 0};
  for (int p = p_begin + threadIdx.x; p < p_end; p += blockDim.x) {
    for (int c = 0; c < C; ++c) {
      float q = __fadd_rn(__fmul_rn(img[(b * HW + p) * C + c], 127.5f), 128.0f);
      q = fminf(fmaxf(q, 0.0f), 255.0f);
      const uint8_t t = static_cast<uint8_t>(q);
      local[c] += t;
      if (u8 != nullptr) u8[(b * C + c) * HW + p] = t;
    }
  }
  for (int c = 0; c < C; ++c) {
    uint32_t t = local[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(&chan_sums[b * 4 + c], t);      // integer atomics: order independent
  }
}

}  // namespace b200
