// Small kernels around the ADM-64 classifier scorer (EncoderUNetModel, edm/unet.py:701-912): input
// scaling, the CLIP-style attention pool (edm/unet.py:40-69) restricted to the only token whose
// output is used (token 0), and softmax + target-probability gather (edm/scorers.py:162-172).
// The torso (ResBlocks / AttentionBlocks) runs on the U-Net engine's tcgen05 kernels.
#pragma once
#include "common.cuh"

namespace b200 {

// images.float() / 255.0                                              (edm/scorers.py:153)
__global__ void u8_to_unit_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __fdiv_rn(static_cast<float>(in[i]), 255.0f);
}

// x = cat([mean(x), x]) + positional_embedding                       (edm/unet.py:63-65)
// act [B, T, C] bf16 (T spatial tokens), pos fp32 [C, T+1] -> tok [B, T, C] bf16 (spatial tokens),
// tok0 [B, C] fp32 (the mean token).  grid (B), any block size.
__global__ void pool_tokens_kernel(const act_t* __restrict__ act, const float* __restrict__ pos,
                                   act_t* __restrict__ tok, float* __restrict__ tok0, int T, int C) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sum = 0.f;
    for (int t = 0; t < T; ++t) {
      const float v = act2f(act[(static_cast<size_t>(b) * T + t) * C + c]);
      sum += v;
      tok[(static_cast<size_t>(b) * T + t) * C + c] = f2act(v + pos[static_cast<size_t>(c) * (T + 1) + 1 + t]);
    }
    tok0[static_cast<size_t>(b) * C + c] = sum / static_cast<float>(T) + pos[static_cast<size_t>(c) * (T + 1)];
  }
}

// QKVAttention (edm/unet.py:388-407) for query token 0 only: per (sample, head of 64 channels)
//   w = softmax_t( (q0 . k_t) / sqrt(64) ), a = sum_t w_t v_t  over T+1 tokens (token 0 = mean token).
// qkv0 fp32 [B, 3C] = qkv_proj(mean token) as [q | k | v]; kv bf16 [B, T, 2C] = [k | v] of the spatial tokens.
// grid (heads, B), 128 threads, T <= 127.
__global__ void pool_attention_kernel(const float* __restrict__ qkv0, const act_t* __restrict__ kv,
                                      float* __restrict__ out, int T, int C) {
  __shared__ float s_q[64];
  __shared__ float s_w[128];
  __shared__ float s_red[2];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* q0 = qkv0 + static_cast<size_t>(b) * 3 * C;
  if (tid < 64) s_q[tid] = q0[h * 64 + tid];
  __syncthreads();
  float s = -INFINITY;
  if (tid <= T) {
    float acc = 0.f;
    if (tid == 0) {
      for (int d = 0; d < 64; ++d) acc = fmaf(s_q[d], q0[C + h * 64 + d], acc);
    } else {
      const act_t* kp = kv + (static_cast<size_t>(b) * T + (tid - 1)) * 2 * C + h * 64;
      for (int d = 0; d < 64; ++d) acc = fmaf(s_q[d], act2f(kp[d]), acc);
    }
    s = acc * 0.125f;
  }
  s_w[tid] = s;
  __syncthreads();
  if (tid == 0) {
    float mx = -INFINITY;
    for (int t = 0; t <= T; ++t) mx = fmaxf(mx, s_w[t]);
    s_red[0] = mx;
  }
  __syncthreads();
  const float e = tid <= T ? expf(s - s_red[0]) : 0.f;
  s_w[tid] = e;
  __syncthreads();
  if (tid == 0) {
    float sum = 0.f;
    for (int t = 0; t <= T; ++t) sum += s_w[t];
    s_red[1] = sum;
  }
  __syncthreads();
  if (tid < 64) {
    float acc = s_w[0] * q0[2 * C + h * 64 + tid];
    for (int t = 0; t < T; ++t)
      acc = fmaf(s_w[t + 1], act2f(kv[(static_cast<size_t>(b) * T + t) * 2 * C + C + h * 64 + tid]), acc);
    out[static_cast<size_t>(b) * C + h * 64 + tid] = acc / s_red[1];
  }
}

// scores[r] = softmax(logits[r, :])[target[r]]                        (edm/scorers.py:163-172)
__global__ void softmax_gather_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                                      float* __restrict__ scores, int K) {
  __shared__ float s_part[32];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* lp = logits + static_cast<size_t>(r) * K;
  float mx = -INFINITY;
  for (int k = tid; k < K; k += blockDim.x) mx = fmaxf(mx, lp[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_part[tid >> 5] = mx;
  __syncthreads();
  mx = s_part[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, s_part[w]);
  __syncthreads();
  float sum = 0.f;
  for (int k = tid; k < K; k += blockDim.x) sum += expf(lp[k] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((tid & 31) == 0) s_part[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += s_part[w];
    scores[r] = expf(lp[target[r]] - mx) / tot;
  }
}

}  // namespace b200
