// HBM-bound leaves of the SD-1.5-shaped UNet2DConditionModel (BASELINE.json config 5) and the DDIM beam step:
//   layernorm_kernel   nn.LayerNorm over channels of a token tensor      (attention.py BasicTransformerBlock norm1-3)
//   geglu_kernel       hidden * gelu(gate), erf GELU (1.5e-7)              (activations.py GEGLU.forward :117-123)
//   upsample2x_kernel  F.interpolate(scale_factor=2, mode='nearest')     (upsampling.py Upsample2D.forward)
//   ddim_cfg_step_kernel / ddim_x0_score_kernel                          (pipeline_stable_diffusion.py:1073-1123,
//                                                                         scheduling_ddim.py:398-460, sd/scorers.py:66-67)
// The sampler arithmetic is fp32 with explicit round-to-nearest intrinsics in the reference's op order (no FMA
// contraction), so it is bit-exact against torch given the same network output.
#pragma once
#include "common.cuh"
#include "groupnorm.cuh"

namespace b200 {

// LPR lanes per token row (a warp normalises 32/LPR consecutive rows at once), each lane holding up to CPL 16-byte
// chunks of its row in registers: all loads of a warp are issued before the first reduction, so a C=320 row (40 chunks,
// LPR=8, 5 chunks per lane) keeps 4 rows x 640 B in flight per warp instead of one.  C multiple of 8, C/8 <= LPR*CPL;
// bf16 in/out, fp32 statistics (two-pass over the registers, fixed shuffle order => batch-position invariant).
template <int LPR, int CPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const act_t* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, act_t* __restrict__ out,
                                                        int64_t rows, int C, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int RPW = 32 / LPR;                   // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t row = warp * RPW + lane / LPR;
  const bool live = row < rows;
  const act_t* xr = x + (live ? row : 0) * C;
  const int nchunk = C >> 3;
  float v[CPL][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int ch = sub + i * LPR;
    if (ch < nchunk) load8(xr + ch * 8, v[i]);
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int ch = sub + i * LPR;
    if (ch < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int ch = sub + i * LPR;
    if (ch < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / static_cast<float>(C) + eps);
  if (!live) return;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int ch = sub + i * LPR;
    if (ch < nchunk) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8) + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
      store8(out + row * C + ch * 8, o8);
    }
  }
}

// in [rows, 2F] = [hidden | gate] -> out [rows, F] = hidden * gelu(gate)   (F multiple of 8)
__global__ void __launch_bounds__(256) geglu_kernel(const act_t* __restrict__ in, act_t* __restrict__ out,
                                                    int64_t rows, int F) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = rows * (F >> 3);
  const int fc = F >> 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / fc;
    const int c = static_cast<int>(i - r * fc) * 8;
    float a[8], g[8], o[8];
    load8(in + r * 2 * F + c, a);
    load8(in + r * 2 * F + F + c, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] * gelu_erf(g[j]);
    store8(out + r * F + c, o);
  }
}

// bf16 NHWC [B,H,W,C] -> [B,2H,2W,C], nearest
__global__ void __launch_bounds__(256) upsample2x_kernel(const act_t* __restrict__ in, act_t* __restrict__ out,
                                                         int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int cc = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * 2 * H * 2 * W * cc;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cc);
    int64_t p = i / cc;
    const int ox = static_cast<int>(p % (2 * W));
    p /= 2 * W;
    const int oy = static_cast<int>(p % (2 * H));
    const int b = static_cast<int>(p / (2 * H));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<int64_t>(b) * H + (oy >> 1)) * W + (ox >> 1)) * C) + c);
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// Classifier-free guidance + DDIM step with variance noise for R candidate rows (each row r belongs to parent
// p = r / per_parent):  eps = eu + g*(et - eu)                                   pipeline...:1073-1075
//   x0 = (sample - sqrt(1-a_t)*eps) / sqrt(a_t); prev = sqrt(a_prev)*x0 + dir_coef*eps + std*noise   scheduling_ddim.py:398-460
// eps_u / eps_t: UNet output fp32 NHWC [P, H, W, C]; sample fp32 NCHW [P, C, H, W]; noise fp32 NCHW [R, C, H, W].
// Writes prev (fp32 NCHW [R, ...]) and, when net_in != null, the next UNet input for both CFG halves
// (net_in[r] = net_in[R + r] = prev[r]; scale_model_input is the identity for DDIM).
__global__ void __launch_bounds__(256) ddim_cfg_step_kernel(const float* __restrict__ eps_u, const float* __restrict__ eps_t,
                                                            const float* __restrict__ sample, const float* __restrict__ noise,
                                                            float* __restrict__ prev, float* __restrict__ net_in, int64_t R,
                                                            int per_parent, int C, int HW, float guidance, float sqrt_beta_t,
                                                            float sqrt_alpha_t, float sqrt_alpha_prev, float dir_coef, float std) {
  const int64_t E = static_cast<int64_t>(C) * HW;
  const int64_t total = R * E;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / E;
    const int64_t e = i - r * E;
    const int c = static_cast<int>(e / HW), px = static_cast<int>(e - static_cast<int64_t>(c) * HW);
    const int64_t p = r / per_parent;
    const int64_t nhwc = (p * HW + px) * C + c;
    const float eu = eps_u[nhwc], et = eps_t[nhwc];
    const float eps = __fadd_rn(eu, __fmul_rn(guidance, __fsub_rn(et, eu)));
    const float x = sample[p * E + e];
    const float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(sqrt_beta_t, eps)), sqrt_alpha_t);
    float pv = __fadd_rn(__fmul_rn(sqrt_alpha_prev, x0), __fmul_rn(dir_coef, eps));
    if (noise != nullptr) pv = __fadd_rn(pv, __fmul_rn(std, noise[i]));
    prev[i] = pv;
    if (net_in != nullptr) {
      net_in[i] = pv;
      net_in[total + i] = pv;
    }
  }
}

// Guided eps of the second UNet call -> pred_x0 of the candidate -> uint8 quantisation -> mean(u8/255) over (C,H,W)
// (pipeline...:1101-1123 with identity decode; sd/scorers.py:66-67 non-RGB branch).  One CTA per candidate row;
// integer sum (exact), score = sum / (255 * C*H*W) evaluated as fp32(sum/255 mean) like the reference's
// `(u8.float()/255).mean()` up to the fp32 reduction order (<= 1 ulp).  Also writes pred_x0 when asked.
__global__ void __launch_bounds__(256) ddim_x0_score_kernel(const float* __restrict__ eps_u, const float* __restrict__ eps_t,
                                                            const float* __restrict__ cand, float* __restrict__ pred_x0,
                                                            int32_t* __restrict__ sums, float* __restrict__ scores, int C, int HW,
                                                            float guidance, float sqrt_beta_t, float sqrt_alpha_t) {
  __shared__ int s_part[8];
  const int64_t r = blockIdx.x;
  const int64_t E = static_cast<int64_t>(C) * HW;
  int acc = 0;
  for (int64_t e = threadIdx.x; e < E; e += blockDim.x) {
    const int c = static_cast<int>(e / HW), px = static_cast<int>(e - static_cast<int64_t>(c) * HW);
    const int64_t nhwc = (r * HW + px) * C + c;
    const float eu = eps_u[nhwc], et = eps_t[nhwc];
    const float eps = __fadd_rn(eu, __fmul_rn(guidance, __fsub_rn(et, eu)));
    const float x0 = __fdiv_rn(__fsub_rn(cand[r * E + e], __fmul_rn(sqrt_beta_t, eps)), sqrt_alpha_t);
    if (pred_x0 != nullptr) pred_x0[r * E + e] = x0;
    float q = __fadd_rn(__fmul_rn(x0, 127.5f), 128.0f);
    q = fminf(fmaxf(q, 0.0f), 255.0f);
    acc += static_cast<int>(static_cast<uint8_t>(q));          // truncating cast, as .to(torch.uint8)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += s_part[w];
    if (sums != nullptr) sums[r] = t;
    scores[r] = static_cast<float>(static_cast<double>(t) / (255.0 * static_cast<double>(E)));
  }
}

// SD eps_greedy / zero_order candidate noises (pipeline_stable_diffusion.py:1368-1378), fp32, one CTA per candidate n:
//   fresh[n] != 0 :  cand = dirs[n]                                            (a fresh N(0,I) draw, :1375)
//   else          :  cand = pivot + (((dirs[n] / ||dirs[n]||_2) * u[n]) * lambda) * sqrt(C*H*W)   (:1377-1379, the reference's
//                    left-to-right tensor-scalar products, each rounded to fp32)
// The norm is a fixed-order fp32 block reduction (torch.norm's summation order is not reproducible outside ATen: the
// candidates agree with the reference's to fp32 rounding, identical inputs always give identical bits).
__global__ void __launch_bounds__(256) sd_candidates_kernel(const float* __restrict__ pivot, const float* __restrict__ dirs,
                                                            const float* __restrict__ u, const uint8_t* __restrict__ fresh,
                                                            float* __restrict__ cand, int64_t E, float lambda, float sqrt_e) {
  __shared__ float s_part[8];
  __shared__ float s_norm;
  const int64_t n = blockIdx.x;
  const float* d = dirs + n * E;
  float* o = cand + n * E;
  if (fresh != nullptr && fresh[n] != 0) {
    for (int64_t e = threadIdx.x; e < E; e += blockDim.x) o[e] = d[e];
    return;
  }
  float acc = 0.f;
  for (int64_t e = threadIdx.x; e < E; e += blockDim.x) acc = fmaf(d[e], d[e], acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += s_part[w];
    s_norm = sqrtf(t);
  }
  __syncthreads();
  const float nrm = s_norm, un = u[n];
  for (int64_t e = threadIdx.x; e < E; e += blockDim.x)
    o[e] = __fadd_rn(pivot[e], __fmul_rn(__fmul_rn(__fmul_rn(__fdiv_rn(d[e], nrm), un), lambda), sqrt_e));
}

}  // namespace b200
