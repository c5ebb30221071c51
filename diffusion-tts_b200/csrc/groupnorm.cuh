// GroupNorm for bf16 NHWC activations, split in two HBM-bound passes:
//   gn_stats_kernel : per-(sample, split, group) partial sum / sum-of-squares (fp64 partials,
//                     fixed reduction order -> bit-identical for identical inputs regardless
//                     of batch position or shard, which exact-tie argmax relies on)
//   gn_apply_kernel : finalise mean/rstd, then y = act(FiLM(norm(x))) with optional 2x resample
// Both read up to two concatenated sources (decoder skip concat, networks.py:458) so the
// concat is never materialised.  Replaces GroupNorm.forward (edm/training/networks.py:104-106)
// and the silu/addcmul glue of UNetBlock.forward (:168, :173-175, :182).
#pragma once
#include "common.cuh"

namespace b200 {

struct GnStatsArgs {
  const act_t* x0;
  const act_t* x1;
  int C0, C1, C;           // C = C0 + C1
  int HW, groups, cpg;
  const float* pre_add;    // [b_emb, ld_pre_add] or null
  int ld_pre_add, b_emb;
  double* partial;         // [batch, splits, groups, 2]
  int splits, PY;
};

DEVINL void load8(const act_t* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  float2 t;
  t = unpack_act(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_act(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_act(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_act(u.w); f[6] = t.x; f[7] = t.y;
}
DEVINL uint4 load_raw(const act_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
DEVINL void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_act(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_act(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_act(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_act(u.w); f[6] = t.x; f[7] = t.y;
}
DEVINL void store8(act_t* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_act(f[0], f[1]);
  u.y = pack_act(f[2], f[3]);
  u.z = pack_act(f[4], f[5]);
  u.w = pack_act(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// grid (splits, batch); block = (C/8) * PY threads (<= 256)
__global__ void __launch_bounds__(256, 4) gn_stats_kernel(const GnStatsArgs a) {
  __shared__ float s_sum[2048];
  __shared__ float s_sq[2048];
  const int VC = a.C >> 3;
  const int vx = threadIdx.x % VC, py = threadIdx.x / VC;
  const int split = blockIdx.x, bi = blockIdx.y;
  const int ppb = a.HW / a.splits;
  const int c = vx * 8;
  const act_t* src;
  int ld;
  if (c < a.C0) {
    src = a.x0 + c;
    ld = a.C0;
  } else {
    src = a.x1 + (c - a.C0);
    ld = a.C1;
  }
  src += (static_cast<size_t>(bi) * a.HW + static_cast<size_t>(split) * ppb) * ld;
  float pa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (a.pre_add != nullptr) {
    const float* pp = a.pre_add + static_cast<size_t>(bi % a.b_emb) * a.ld_pre_add + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) pa[j] = pp[j];
  }
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ss[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // 4 independent 16-byte loads in flight per thread; the pixel order per thread is fixed
  // (p = py, py+PY, ...), so the sums do not depend on batch size or position.
  int p = py;
  for (; p + 7 * a.PY < ppb; p += 8 * a.PY) {
    uint4 raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) raw[u] = load_raw(src + static_cast<size_t>(p + u * a.PY) * ld);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float f[8];
      unpack8(raw[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = f[j] + pa[j];
        s[j] += v;
        ss[j] += v * v;
      }
    }
  }
  for (; p < ppb; p += a.PY) {
    float f[8];
    load8(src + static_cast<size_t>(p) * ld, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = f[j] + pa[j];
      s[j] += v;
      ss[j] += v * v;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_sum[py * a.C + c + j] = s[j];
    s_sq[py * a.C + c + j] = ss[j];
  }
  __syncthreads();
  // level 1 (all threads): per-channel sum over py in fixed order; level 2: per-group fp64 sum
  float cs[11], cq[11];          // ceil(2048 / 192) channels per thread at most
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int cc = threadIdx.x + k * blockDim.x;
    float ts = 0.f, tq = 0.f;
    if (cc < a.C)
      for (int y = 0; y < a.PY; ++y) {
        ts += s_sum[y * a.C + cc];
        tq += s_sq[y * a.C + cc];
      }
    cs[k] = ts;
    cq[k] = tq;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const int cc = threadIdx.x + k * blockDim.x;
    if (cc < a.C) {
      s_sum[cc] = cs[k];
      s_sq[cc] = cq[k];
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < a.groups; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int cc = g * a.cpg; cc < (g + 1) * a.cpg; ++cc) {
      ds += static_cast<double>(s_sum[cc]);
      dq += static_cast<double>(s_sq[cc]);
    }
    double* o = a.partial + ((static_cast<size_t>(bi) * a.splits + split) * a.groups + g) * 2;
    o[0] = ds;
    o[1] = dq;
  }
}

// GroupNorm statistics from the per-channel (sum, sum of squares) the producing GEMMs wrote per 64-row half
// tile (gemm_conv.cuh): fp64 accumulation, fixed order -> independent of batch size / position / shard.
struct GnFinalizeArgs {
  const float2* st0;
  const float2* st1;
  int C0, C1, C;
  int HW, groups, cpg;
  const float* pre_add;
  int ld_pre_add, b_emb;
  float eps;
  float2* mean_rstd;       // [batch, groups]
};

// WIDE = false: one warp per (sample, group); 8 warps per block; grid = ceil(batch*groups / 8).
// WIDE = true : one 8-warp CTA per (sample, group) -- each warp reduces a contiguous eighth of the (row, channel) list
//               and the eight partial sums are combined in warp order: for the VAE decoder's 512x512 activations
//               (4096 statistics rows per sample) a single warp per pair left the GPU idle (271 us per launch).
// Which variant runs depends only on the tensor shape, so results stay batch-size and batch-position invariant.
// (sum, sum of squares) of one (sample, group) over elements [begin, total) of its (statistics row, channel) list, by ONE
// warp: lane-strided fp64 accumulation with 4 loads in flight, then a butterfly -- every lane ends with the same, order-fixed
// pair.  Shared by gn_finalize_kernel and gn_norm_cluster_kernel, so both produce the same bits.
DEVINL void gn_group_sums(const GnFinalizeArgs& a, int bi, int g, int begin, int total, int lane, double& ds, double& dq) {
  const int R = a.HW >> 6;
  ds = 0.0;
  dq = 0.0;
  auto fetch = [&](int i, float2& v, double& pa) {
    const int r = i / a.cpg;
    const int cc = g * a.cpg + (i - r * a.cpg);
    const size_t prow = static_cast<size_t>(bi) * R + r;
    v = cc < a.C0 ? __ldg(a.st0 + prow * a.C0 + cc) : __ldg(a.st1 + prow * a.C1 + (cc - a.C0));
    pa = a.pre_add != nullptr ? static_cast<double>(a.pre_add[static_cast<size_t>(bi % a.b_emb) * a.ld_pre_add + cc]) : 0.0;
  };
  auto accum = [&](const float2& v, double pa) {
    const double s = static_cast<double>(v.x), q = static_cast<double>(v.y);
    dq += q + 2.0 * pa * s + 64.0 * pa * pa;
    ds += s + 64.0 * pa;
  };
  int i = begin + lane;
  for (; i + 3 * 32 < total; i += 4 * 32) {       // 4 loads in flight; accumulation order stays fixed
    float2 v0, v1, v2, v3;
    double p0, p1, p2, p3;
    fetch(i, v0, p0);
    fetch(i + 32, v1, p1);
    fetch(i + 64, v2, p2);
    fetch(i + 96, v3, p3);
    accum(v0, p0);
    accum(v1, p1);
    accum(v2, p2);
    accum(v3, p3);
  }
  for (; i < total; i += 32) {
    float2 v;
    double pa;
    fetch(i, v, pa);
    accum(v, pa);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {               // butterfly: every lane ends with the same, order-fixed sum
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dq += __shfl_xor_sync(0xffffffffu, dq, o);
  }
}
DEVINL float2 gn_mean_rstd(const GnFinalizeArgs& a, double ds, double dq) {
  const double n = static_cast<double>(a.HW) * a.cpg;
  const double mean = ds / n;
  double var = dq / n - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  return make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps))));
}

template <bool WIDE>
__global__ void __launch_bounds__(256) gn_finalize_kernel(const GnFinalizeArgs a, int n_pairs) {
  __shared__ double s_part[16];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int pair = WIDE ? blockIdx.x : blockIdx.x * 8 + wid;
  if (pair >= n_pairs) return;
  const int bi = pair / a.groups, g = pair - bi * a.groups;
  const int R = a.HW >> 6;
  const int total_all = R * a.cpg;
  const int chunk = WIDE ? (total_all + 7) / 8 : total_all;
  const int begin = WIDE ? wid * chunk : 0;
  const int total = WIDE ? min(total_all, begin + chunk) : total_all;
  double ds, dq;
  gn_group_sums(a, bi, g, begin, total, lane, ds, dq);
  if (WIDE) {
    if (lane == 0) {
      s_part[2 * wid] = ds;
      s_part[2 * wid + 1] = dq;
    }
    __syncthreads();
    if (wid != 0) return;
    ds = 0.0;
    dq = 0.0;
    for (int w = 0; w < 8; ++w) {                    // fixed order
      ds += s_part[2 * w];
      dq += s_part[2 * w + 1];
    }
  }
  if (lane == 0) a.mean_rstd[static_cast<size_t>(bi) * a.groups + g] = gn_mean_rstd(a, ds, dq);
}

struct GnApplyArgs {
  const act_t* x0;
  const act_t* x1;
  int C0, C1, C;
  int H, W, groups, cpg;   // INPUT spatial dims
  const double* partial;
  int splits;
  float eps;
  const float* gamma;
  const float* beta;
  const float* pre_add;
  int ld_pre_add;
  const float* film_scale;
  const float* film_shift;
  int ld_film, b_emb;
  int silu, resample;      // 0 none, 1 up x2, 2 down x2
  act_t* out;
  act_t* raw_out;
  int PY, ITER;
  const float2* mean_rstd; // [batch, groups] from gn_finalize_kernel (then `partial` is unused)
  int reverse;             // walk samples / pixel chunks last-to-first (L2 hits on what the producer wrote last)
};

// grid (pixel chunks, batch); block = (C/8) * PY threads
// WIDE: more than 2048 channels (SD-1.5 decoder concat 2560): up to 512 threads per block
// y = act(FiLM(norm(x))) for the pixel chunk `bx` of sample `bi`; s_mean / s_rstd hold the sample's group statistics
DEVINL void gn_apply_body(const GnApplyArgs& a, int bi, int bx, const float* s_mean, const float* s_rstd) {
  const int VC = a.C >> 3;
  const int vx = threadIdx.x % VC, py = threadIdx.x / VC;
  const int HW = a.H * a.W;
  const int c = vx * 8;
  const act_t* src;
  int ld;
  if (c < a.C0) {
    src = a.x0 + c;
    ld = a.C0;
  } else {
    src = a.x1 + (c - a.C0);
    ld = a.C1;
  }
  src += static_cast<size_t>(bi) * HW * ld;

  // y = act(x * ka + kb) per channel
  float ka[8], kb[8];
  {
    const int e = bi % a.b_emb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cc = c + j;
      const int g = cc / a.cpg;
      const float rs = s_rstd[g] * a.gamma[cc];
      float k1 = rs;
      float k0 = a.beta[cc] - s_mean[g] * rs;
      if (a.pre_add != nullptr) k0 += a.pre_add[static_cast<size_t>(e) * a.ld_pre_add + cc] * rs;
      if (a.film_scale != nullptr) {
        const float sc = a.film_scale[static_cast<size_t>(e) * a.ld_film + cc] + 1.0f;
        const float sh = a.film_shift[static_cast<size_t>(e) * a.ld_film + cc];
        k1 *= sc;
        k0 = k0 * sc + sh;
      }
      ka[j] = k1;
      kb[j] = k0;
    }
  }

  const int outW = a.resample == 1 ? a.W * 2 : (a.resample == 2 ? a.W / 2 : a.W);
  const int outH = a.resample == 1 ? a.H * 2 : (a.resample == 2 ? a.H / 2 : a.H);
  const size_t out_base = static_cast<size_t>(bi) * outH * outW * a.C + c;
  const int dom = a.resample == 2 ? outH * outW : HW;      // loop domain
  const int p_begin = bx * a.PY * a.ITER + py;

  if (a.resample == 2) {
    for (int it = 0; it < a.ITER; ++it) {
      const int p = p_begin + it * a.PY;
      if (p >= dom) break;
      const int oy = p / outW, ox = p - oy * outW;
      float f[4][8];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        load8(src + static_cast<size_t>((2 * oy + (t >> 1)) * a.W + 2 * ox + (t & 1)) * ld, f[t]);
      float accv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, accr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = f[t][j] * ka[j] + kb[j];
          if (a.silu) v = silu_tanh(v);
          accv[j] += v;
          accr[j] += f[t][j];
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        accv[j] *= 0.25f;
        accr[j] *= 0.25f;
      }
      store8(a.out + out_base + static_cast<size_t>(p) * a.C, accv);
      if (a.raw_out != nullptr) store8(a.raw_out + out_base + static_cast<size_t>(p) * a.C, accr);
    }
  } else {
    for (int it0 = 0; it0 < a.ITER; it0 += 8) {
      uint4 raw[8];                                   // 8 independent 16-byte loads in flight per thread
      int pp[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        pp[u] = (it0 + u < a.ITER) ? p_begin + (it0 + u) * a.PY : dom;
        if (pp[u] < dom) raw[u] = load_raw(src + static_cast<size_t>(pp[u]) * ld);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (pp[u] >= dom) continue;
        float f[8], v[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = f[j] * ka[j] + kb[j];
          if (a.silu) v[j] = silu_tanh(v[j]);
        }
        if (a.resample == 0) {
          store8(a.out + out_base + static_cast<size_t>(pp[u]) * a.C, v);
          if (a.raw_out != nullptr)
            *reinterpret_cast<uint4*>(a.raw_out + out_base + static_cast<size_t>(pp[u]) * a.C) = raw[u];
        } else {
          const int iy = pp[u] / a.W, ix = pp[u] - iy * a.W;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const size_t op = static_cast<size_t>((2 * iy + (t >> 1)) * outW + 2 * ix + (t & 1)) * a.C;
            store8(a.out + out_base + op, v);
            if (a.raw_out != nullptr) *reinterpret_cast<uint4*>(a.raw_out + out_base + op) = raw[u];
          }
        }
      }
    }
  }
}

template <bool WIDE>
__global__ void __launch_bounds__(WIDE ? 512 : 256, WIDE ? 1 : 3) gn_apply_kernel(const GnApplyArgs a) {
  __shared__ float s_mean[64];
  __shared__ float s_rstd[64];
  pdl_launch_dependents();
  pdl_wait();
  const int bi = a.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int bx = a.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int HW = a.H * a.W;
  if (a.mean_rstd != nullptr) {
    if (threadIdx.x < a.groups) {
      const float2 mr = a.mean_rstd[static_cast<size_t>(bi) * a.groups + threadIdx.x];
      s_mean[threadIdx.x] = mr.x;
      s_rstd[threadIdx.x] = mr.y;
    }
  } else if (threadIdx.x < a.groups) {
    const int g = threadIdx.x;
    double ds = 0.0, dq = 0.0;
    for (int sp = 0; sp < a.splits; ++sp) {
      const double* o = a.partial + ((static_cast<size_t>(bi) * a.splits + sp) * a.groups + g) * 2;
      ds += o[0];
      dq += o[1];
    }
    const double n = static_cast<double>(HW) * a.cpg;
    const double mean = ds / n;
    double var = dq / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
  __syncthreads();
  gn_apply_body(a, bi, bx, s_mean, s_rstd);
}

// gn_finalize + gn_apply in ONE launch for the low-resolution levels (H*W <= 256), where both are latency-bound (6-8 us of
// finalize and 13-18 us of apply for 6-19 MB tensors): one thread-block cluster of GN_CLUSTER CTAs per sample.  CTA r of the
// cluster reduces the statistics of groups [r*gpc, (r+1)*gpc) -- one warp per group, the very code of gn_finalize_kernel, so
// the bits are the same -- and writes (mean, rstd) to global memory; after a cluster barrier (release / acquire) every CTA
// reads the sample's groups back from L2 and normalises its own eighth of the pixels.
constexpr int GN_CLUSTER = 8;
__global__ void __launch_bounds__(256, 3) gn_norm_cluster_kernel(const GnApplyArgs a, const GnFinalizeArgs f) {
  __shared__ float s_mean[64];
  __shared__ float s_rstd[64];
  pdl_launch_dependents();
  pdl_wait();
  const int bi = a.reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int rank = blockIdx.x;                          // gridDim.x == GN_CLUSTER == the cluster's x extent
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nfull = blockDim.x >> 5;
  const int gpc = (f.groups + GN_CLUSTER - 1) / GN_CLUSTER;
  if (wid < nfull) {                                    // only complete warps reduce (the butterfly needs 32 lanes)
    const int g_end = min(f.groups, (rank + 1) * gpc);
    const int total = (f.HW >> 6) * f.cpg;
    for (int g = rank * gpc + wid; g < g_end; g += nfull) {
      double ds, dq;
      gn_group_sums(f, bi, g, 0, total, lane, ds, dq);
      if (lane == 0) f.mean_rstd[static_cast<size_t>(bi) * f.groups + g] = gn_mean_rstd(f, ds, dq);
    }
  }
  __threadfence();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (threadIdx.x < a.groups) {
    const float2 mr = __ldcg(f.mean_rstd + static_cast<size_t>(bi) * f.groups + threadIdx.x);
    s_mean[threadIdx.x] = mr.x;
    s_rstd[threadIdx.x] = mr.y;
  }
  __syncthreads();
  gn_apply_body(a, bi, a.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x, s_mean, s_rstd);
}

}  // namespace b200
