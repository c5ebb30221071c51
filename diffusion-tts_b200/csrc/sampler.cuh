// Fused sampler / scorer kernels (HBM-bound, fp64 state like the reference).
//
// Bit-faithfulness (SURVEY.md 8a'): PyTorch evaluates edm/main.py:85-94 as one rounded
// elementwise kernel per binary op, so every intermediate below is rounded separately with
// __dmul_rn/__dadd_rn/__dsub_rn/__ddiv_rn (and __fmul_rn/__fadd_rn for the fp32
// preconditioning of networks.py:667) -- no FMA contraction.  Fed an identical network output
// these kernels reproduce the reference's x_hat / x_next / denoised / uint8 image bit for bit.
#pragma once
#include <climits>

#include "common.cuh"

namespace b200 {

// x_hat = x_cur + s*eps ; net_in = c_in * fp32(x_hat)        (edm/main.py:85; networks.py:655,665)
// NoiseT = double: eps is fp64 like x_cur (randn_like(x_cur), :750-800) and s*eps is an fp64 product.
// NoiseT = float : eps is an fp32 tensor (the MCTS depth noises, :445, or an fp32 precomputed_noise): torch multiplies the
//                  0-dim fp64 scale into it IN fp32 (the scalar is rounded to fp32 first) and only the sum is fp64.
template <typename NoiseT>
__global__ void heun_pre_kernel(const double* __restrict__ x_cur, const NoiseT* __restrict__ eps,
                                double* __restrict__ x_hat, float* __restrict__ net_in, int64_t total,
                                int64_t bE, double s, float c_in) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 2;
  const float s32 = static_cast<float>(s);
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < total; i += stride) {
    const double2 x = *reinterpret_cast<const double2*>(x_cur + (i % bE));
    double2 h;
    if constexpr (sizeof(NoiseT) == 4) {
      const float2 e = *reinterpret_cast<const float2*>(eps + i);
      h.x = __dadd_rn(x.x, static_cast<double>(__fmul_rn(s32, e.x)));
      h.y = __dadd_rn(x.y, static_cast<double>(__fmul_rn(s32, e.y)));
    } else {
      const double2 e = *reinterpret_cast<const double2*>(eps + i);
      h.x = __dadd_rn(x.x, __dmul_rn(s, e.x));
      h.y = __dadd_rn(x.y, __dmul_rn(s, e.y));
    }
    *reinterpret_cast<double2*>(x_hat + i) = h;
    float2 o;
    o.x = __fmul_rn(c_in, static_cast<float>(h.x));
    o.y = __fmul_rn(c_in, static_cast<float>(h.y));
    *reinterpret_cast<float2*>(net_in + i) = o;
  }
}

struct HeunCoef {
  float c_skip1, c_out1;
  double t_hat, dt;
  float c_skip2, c_out2;
  double t_next;
  float c_in_next;
};

DEVINL double denoise(double xh, float F, float c_skip, float c_out) {
  const float x32 = static_cast<float>(xh);
  return static_cast<double>(__fadd_rn(__fmul_rn(c_skip, x32), __fmul_rn(c_out, F)));
}

// Euler half step.  F1 is NHWC [R,HW,C]; state is NCHW [R,C,HW].
__global__ void heun_mid_kernel(const double* __restrict__ x_hat, const float* __restrict__ F1,
                                float* __restrict__ net_in2, double* __restrict__ x_eul, int64_t total, int C,
                                int HW, HeunCoef k) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int E = C * HW;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / E;
    const int e = static_cast<int>(i - r * E);
    const int ch = e / HW, p = e - ch * HW;
    const double xh = x_hat[i];
    const double D1 = denoise(xh, F1[(r * HW + p) * C + ch], k.c_skip1, k.c_out1);
    const double d_cur = __ddiv_rn(__dsub_rn(xh, D1), k.t_hat);
    const double xe = __dadd_rn(xh, __dmul_rn(k.dt, d_cur));
    if (x_eul != nullptr) x_eul[i] = xe;
    net_in2[i] = __fmul_rn(k.c_in_next, static_cast<float>(xe));
  }
}

// Heun correction + Tweedie x0 + uint8 quantise + integer channel sums.
// grid (chunks, R); each CTA handles a contiguous pixel range of one candidate row.
__global__ void heun_post_kernel(const double* __restrict__ x_hat, const float* __restrict__ F1,
                                 const float* __restrict__ F2, double* __restrict__ x_next,
                                 uint8_t* __restrict__ x0_u8, uint32_t* __restrict__ chan_sums, int C, int HW,
                                 HeunCoef k) {
  __shared__ uint32_t s_sum[4];
  if (threadIdx.x < 4) s_sum[threadIdx.x] = 0;
  __syncthreads();
  const int64_t r = blockIdx.y;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * per, p_end = min(HW, p_begin + per);
  for (int ch = 0; ch < C; ++ch) {
    uint32_t local = 0;
    for (int p = p_begin + threadIdx.x; p < p_end; p += blockDim.x) {
      const int64_t i = (r * C + ch) * HW + p;
      const double xh = x_hat[i];
      const int64_t fi = (r * HW + p) * C + ch;
      const double D1 = denoise(xh, F1[fi], k.c_skip1, k.c_out1);
      const double d_cur = __ddiv_rn(__dsub_rn(xh, D1), k.t_hat);
      double xn = __dadd_rn(xh, __dmul_rn(k.dt, d_cur));
      double den = D1;
      if (F2 != nullptr) {
        den = denoise(xn, F2[fi], k.c_skip2, k.c_out2);
        const double d_prime = __ddiv_rn(__dsub_rn(xn, den), k.t_next);
        const double avg = __dadd_rn(__dmul_rn(0.5, d_cur), __dmul_rn(0.5, d_prime));
        xn = __dadd_rn(xh, __dmul_rn(k.dt, avg));
      }
      if (x_next != nullptr) x_next[i] = xn;
      // (x*127.5+128).clip(0,255).to(uint8)
      double q = __dadd_rn(__dmul_rn(den, 127.5), 128.0);
      q = fmin(fmax(q, 0.0), 255.0);
      const uint32_t u = static_cast<uint32_t>(q);       // truncation; NaN -> 0
      if (x0_u8 != nullptr) x0_u8[i] = static_cast<uint8_t>(u);
      local += u;
    }
    if (chan_sums != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
      if ((threadIdx.x & 31) == 0 && ch < 4) atomicAdd(&s_sum[ch], local);
    }
  }
  if (chan_sums != nullptr) {
    __syncthreads();
    if (threadIdx.x < C && threadIdx.x < 4) atomicAdd(&chan_sums[r * 4 + threadIdx.x], s_sum[threadIdx.x]);
  }
}

// ---- C == 3 fast paths (RGB images, every EDM config): one thread owns TWO adjacent pixels of all three channels, so the
// NHWC network output is read as 24 contiguous bytes per thread (three float2; a warp reads 768 contiguous bytes -- the
// scalar kernels above fetch every 32-byte sector of F three times, once per channel pass) and the planar fp64 state moves
// as double2.  Arithmetic per element is identical to the scalar kernels (same rounded operations in the same order).
DEVINL void heun_elem(double xh, float f1, float f2, bool second, const HeunCoef& k, double& xn, double& den) {
  const double D1 = denoise(xh, f1, k.c_skip1, k.c_out1);
  const double d_cur = __ddiv_rn(__dsub_rn(xh, D1), k.t_hat);
  xn = __dadd_rn(xh, __dmul_rn(k.dt, d_cur));
  den = D1;
  if (second) {
    den = denoise(xn, f2, k.c_skip2, k.c_out2);
    const double d_prime = __ddiv_rn(__dsub_rn(xn, den), k.t_next);
    const double avg = __dadd_rn(__dmul_rn(0.5, d_cur), __dmul_rn(0.5, d_prime));
    xn = __dadd_rn(xh, __dmul_rn(k.dt, avg));
  }
}

// total2 = R * HW / 2 pixel pairs
__global__ void heun_mid_c3_kernel(const double* __restrict__ x_hat, const float* __restrict__ F1, float* __restrict__ net_in2,
                                   double* __restrict__ x_eul, int64_t total2, int HW, HeunCoef k) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int half = HW >> 1;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total2; t += stride) {
    const int64_t r = t / half;
    const int p = static_cast<int>(t - r * half) * 2;
    const float2* fp = reinterpret_cast<const float2*>(F1 + (r * HW + p) * 3);
    const float2 a = __ldg(fp), b = __ldg(fp + 1), c = __ldg(fp + 2);          // (p,0) (p,1) | (p,2) (p+1,0) | (p+1,1) (p+1,2)
    const float f[2][3] = {{a.x, a.y, b.x}, {b.y, c.x, c.y}};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int64_t i = (r * 3 + ch) * HW + p;
      const double2 xh = *reinterpret_cast<const double2*>(x_hat + i);
      double2 xe;
      double den;
      heun_elem(xh.x, f[0][ch], 0.f, false, k, xe.x, den);
      heun_elem(xh.y, f[1][ch], 0.f, false, k, xe.y, den);
      if (x_eul != nullptr) *reinterpret_cast<double2*>(x_eul + i) = xe;
      *reinterpret_cast<float2*>(net_in2 + i) =
          make_float2(__fmul_rn(k.c_in_next, static_cast<float>(xe.x)), __fmul_rn(k.c_in_next, static_cast<float>(xe.y)));
    }
  }
}

// grid (chunks, R); each CTA handles a contiguous range of pixel pairs of one candidate row
__global__ void heun_post_c3_kernel(const double* __restrict__ x_hat, const float* __restrict__ F1, const float* __restrict__ F2,
                                    double* __restrict__ x_next, uint8_t* __restrict__ x0_u8, uint32_t* __restrict__ chan_sums,
                                    int HW, HeunCoef k) {
  __shared__ uint32_t s_sum[3];
  if (threadIdx.x < 3) s_sum[threadIdx.x] = 0;
  __syncthreads();
  const int64_t r = blockIdx.y;
  const int half = HW >> 1;
  const int per = (half + gridDim.x - 1) / gridDim.x;
  const int q_begin = blockIdx.x * per, q_end = min(half, q_begin + per);
  const bool second = F2 != nullptr;
  uint32_t local[3] = {0, 0, 0};
  for (int q = q_begin + threadIdx.x; q < q_end; q += blockDim.x) {
    const int p = 2 * q;
    const float2* fp = reinterpret_cast<const float2*>(F1 + (r * HW + p) * 3);
    const float2 a = __ldg(fp), b = __ldg(fp + 1), c = __ldg(fp + 2);
    const float f1[2][3] = {{a.x, a.y, b.x}, {b.y, c.x, c.y}};
    float f2[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    if (second) {
      const float2* gp = reinterpret_cast<const float2*>(F2 + (r * HW + p) * 3);
      const float2 a2 = __ldg(gp), b2 = __ldg(gp + 1), c2 = __ldg(gp + 2);
      f2[0][0] = a2.x; f2[0][1] = a2.y; f2[0][2] = b2.x;
      f2[1][0] = b2.y; f2[1][1] = c2.x; f2[1][2] = c2.y;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int64_t i = (r * 3 + ch) * HW + p;
      const double2 xh = *reinterpret_cast<const double2*>(x_hat + i);
      double2 xn;
      double d0, d1;
      heun_elem(xh.x, f1[0][ch], f2[0][ch], second, k, xn.x, d0);
      heun_elem(xh.y, f1[1][ch], f2[1][ch], second, k, xn.y, d1);
      if (x_next != nullptr) *reinterpret_cast<double2*>(x_next + i) = xn;
      double q0 = fmin(fmax(__dadd_rn(__dmul_rn(d0, 127.5), 128.0), 0.0), 255.0);      // (x*127.5+128).clip(0,255).to(uint8)
      double q1 = fmin(fmax(__dadd_rn(__dmul_rn(d1, 127.5), 128.0), 0.0), 255.0);
      const uint32_t u0 = static_cast<uint32_t>(q0), u1 = static_cast<uint32_t>(q1);   // truncation; NaN -> 0
      if (x0_u8 != nullptr) *reinterpret_cast<uchar2*>(x0_u8 + i) = make_uchar2(static_cast<uint8_t>(u0), static_cast<uint8_t>(u1));
      local[ch] += u0 + u1;
    }
  }
  if (chan_sums != nullptr) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      uint32_t v = local[ch];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum[ch], v);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(&chan_sums[r * 4 + threadIdx.x], s_sum[threadIdx.x]);
  }
}

__global__ void quantize_u8_kernel(const double* __restrict__ x, uint8_t* __restrict__ out, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    double q = __dadd_rn(__dmul_rn(x[i], 127.5), 128.0);
    q = fmin(fmax(q, 0.0), 255.0);
    out[i] = static_cast<uint8_t>(static_cast<uint32_t>(q));
  }
}

// grid (chunks, M): integer channel sums of uint8 [M,C,HW]
__global__ void channel_sums_u8_kernel(const uint8_t* __restrict__ img, uint32_t* __restrict__ chan_sums, int C,
                                       int HW) {
  const int64_t r = blockIdx.y;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * per, p_end = min(HW, p_begin + per);
  for (int ch = 0; ch < C && ch < 4; ++ch) {
    uint32_t local = 0;
    for (int p = p_begin + threadIdx.x; p < p_end; p += blockDim.x) local += img[(r * C + ch) * HW + p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&chan_sums[r * 4 + ch], local);
  }
}

// BrightnessScorer from exact integer sums (edm/scorers.py:37-52; sd/scorers.py:66-67).
__global__ void brightness_kernel(const uint32_t* __restrict__ chan_sums, float* __restrict__ scores, int64_t M,
                                  int C, int HW) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const uint32_t* s = chan_sums + i * 4;
  double v;
  if (C == 3) {
    v = (0.2126 * s[0] + 0.7152 * s[1] + 0.0722 * s[2]) / (255.0 * HW);
  } else {
    double t = 0.0;
    for (int ch = 0; ch < C && ch < 4; ++ch) t += s[ch];
    v = t / (255.0 * HW * C);
  }
  float f = static_cast<float>(v);
  scores[i] = fminf(fmaxf(f, 0.0f), 1.0f);
}

// Signed-orderable int32 image of a float: a < b  <=>  key(a) < key(b) as int32.
DEVINL int32_t orderable_f32(float f) {
  const int32_t i = __float_as_int(f);
  return i < 0 ? (i ^ 0x7FFFFFFF) : i;
}

// One warp per image j: first-maximal index over n (edm/main.py:842) via a signed-int64 max of
// (orderable(score) << 32) | (0xFFFFFFFF - global_index): the larger score wins, and among equal
// scores the SMALLER index.  The same key reduces across shards with ncclAllReduce(max, int64).
__global__ void argmax_first_kernel(const float* __restrict__ scores, int64_t N, int64_t b, int64_t idx_base,
                                    int64_t* __restrict__ idx, int64_t* __restrict__ packed) {
  const int64_t j = blockIdx.x;
  long long best = LLONG_MIN;
  for (int64_t n = threadIdx.x; n < N; n += 32) {
    float s = scores[n * b + j];
    if (s != s) s = -INFINITY;     // NaN never wins
    const long long key = (static_cast<long long>(orderable_f32(s)) << 32) |
                          static_cast<long long>(0xFFFFFFFFu - static_cast<uint32_t>(idx_base + n));
    best = key > best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if (threadIdx.x == 0) {
    if (idx != nullptr)
      idx[j] = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(best & 0xFFFFFFFFll)) - idx_base;
    if (packed != nullptr) packed[j] = best;
  }
}

// dst[j,:] = src[idx[j], j, :]
__global__ void gather_rows_kernel(const double* __restrict__ src, const int64_t* __restrict__ idx,
                                   double* __restrict__ dst, int64_t b, int64_t E) {
  const int64_t j = blockIdx.y;
  const int64_t n = idx[j];
  const double* s = src + (n * b + j) * E;
  double* d = dst + j * E;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < E;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x)
    d[e] = s[e];
}

// ||dirs[r,:]||_2 -- one CTA per row, fixed-order tree reduction in fp64.
__global__ void direction_norms_kernel(const double* __restrict__ dirs, double* __restrict__ norms, int64_t E) {
  __shared__ double s_part[32];
  const int64_t r = blockIdx.x;
  const double* d = dirs + r * E;
  double acc = 0.0;
  for (int64_t e = threadIdx.x; e < E; e += blockDim.x) {
    const double v = d[e];
    acc = fma(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) norms[r] = sqrt(v);
  }
}

// cand = fresh ? fresh : pivot + (double)scale * (dir / norm)     (edm/main.py:764-795)
__global__ void make_candidates_kernel(const double* __restrict__ pivot, const double* __restrict__ dirs,
                                       const double* __restrict__ norms, const float* __restrict__ scale,
                                       const uint8_t* __restrict__ fresh_mask, const double* __restrict__ fresh,
                                       double* __restrict__ cand, int64_t b, int64_t E) {
  const int64_t r = blockIdx.y;
  const bool is_fresh = fresh_mask != nullptr && fresh_mask[r] != 0;
  const double nrm = norms[r];
  const double sc = static_cast<double>(scale[r]);
  const double* pv = pivot + (r % b) * E;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < E;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = r * E + e;
    if (is_fresh) {
      cand[i] = fresh[i];
    } else {
      const double u = __ddiv_rn(dirs[i], nrm);
      cand[i] = __dadd_rn(pv[e], __dmul_rn(sc, u));
    }
  }
}

// ---------------------------------------------------------------- small U-Net helpers
// out[r,n] = act(sum_k x[r,k] W[n,k] + bias[n] + add[r,n]) -- one warp per (r, n), fp32.
struct LinearArgs {
  const float* x;
  int rows, K, ld_x;
  const float* w;
  const float* bias;
  const float* add;
  int ld_add, N, act;
  float* out;
  int ld_out;
};
__global__ void linear_kernel(const LinearArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_global >= a.rows * a.N) return;
  const int r = warp_global / a.N, n = warp_global - r * a.N;
  const float* xr = a.x + static_cast<size_t>(r) * a.ld_x;
  const float* wr = a.w + static_cast<size_t>(n) * a.K;
  float acc = 0.f;
  for (int k = lane; k < a.K; k += 32) acc = fmaf(xr[k], __ldg(wr + k), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (a.bias != nullptr) acc += a.bias[n];
    if (a.add != nullptr) acc += a.add[static_cast<size_t>(r) * a.ld_add + n];
    if (a.act == 1) acc = acc / (1.0f + expf(-acc));
    a.out[static_cast<size_t>(r) * a.ld_out + n] = acc;
  }
}

// 3x3 im2col of fp32 NCHW [B,C,H,W] (C<=7) -> bf16 [B*H*W, 64]; k = (kh*3+kw)*C + c.
__global__ void im2col_c3_kernel(const float* __restrict__ x, act_t* __restrict__ out, int B, int C, int H,
                                 int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = static_cast<int64_t>(B) * H * W * 8;    // 8 chunks of 8 bf16 per row
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(i & 7);
    const int64_t m = i >> 3;
    const int px = static_cast<int>(m % W);
    const int py = static_cast<int>((m / W) % H);
    const int bi = static_cast<int>(m / (static_cast<int64_t>(W) * H));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = chunk * 8 + j;
      float v = 0.f;
      if (k < 9 * C) {
        const int tap = k / C, c = k - tap * C;
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((static_cast<int64_t>(bi) * C + c) * H + yy) * W + xx];
      }
      f[j] = v;
    }
    uint4 u;
    u.x = pack_act(f[0], f[1]);
    u.y = pack_act(f[2], f[3]);
    u.z = pack_act(f[4], f[5]);
    u.w = pack_act(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + m * 64 + chunk * 8) = u;
  }
}

}  // namespace b200
