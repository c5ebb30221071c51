// fp32-faithful ("precise") variants of the U-Net kernels, used ONLY to re-score the near-tie contenders of a search
// round (SURVEY.md 7 hard part 1, option b): the bf16 tensor-core path carries ~1e-4 of score noise, the reference runs
// the network in fp32 (edm/training/networks.py:655-667), and the selected index must equal the reference's.
//
// Number format "split fp16": a value v is stored as two IEEE half numbers  hi = half(v), lo = half(v - hi)
// (v ~ hi + lo to ~2^-22 relative; absolute floor 2^-25).  An activation tensor [rows, C] becomes [rows, 2C] halves:
// columns [0, C) = hi plane, [C, 2C) = lo plane, so a GEMM over (hi + lo) x (Whi + Wlo) is the same implicit-GEMM main
// loop (TMA boxes of 64 channels, tcgen05.mma kind::f16, fp32 accumulation in TMEM) with three MMAs per K slice:
//     hi x Whi  +  lo x Whi  +  hi x Wlo                  (the lo x Wlo term, 2^-22, is dropped)
// on ONE staged copy of the four operand tiles (hi, lo, Whi, Wlo) per logical K block,
// i.e. 3x the bf16 MMA work instead of the ~30x an FFMA path would cost.  Everything elementwise (GroupNorm, SiLU,
// FiLM, softmax) is evaluated in fp32 / fp64 with IEEE division and expf (no .approx), attention runs on the FMA pipe in
// fp32 (12 GFLOP of the 219 GFLOP per forward).  All reductions are order-fixed: identical inputs give identical bits at
// any batch position, like the bf16 engine.
#pragma once
#include <cuda_fp16.h>

#include "gemm_conv.cuh"

namespace b200 {

// experiment switch (b200ns_debug_prec_nolo): 1 = drop the lo plane, i.e. plain fp16 storage -- measures what a single-plane
// fp16 engine would give against the reference (tools/measure_precise_error.py)
__device__ int g_prec_nolo = 0;

DEVINL void split_h(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = g_prec_nolo ? __float2half_rn(0.f) : __float2half_rn(v - __half2float(hi));
}
DEVINL void unpack8h(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
// 8 consecutive channels of a split tensor: hi at p, lo at p + lo_off
DEVINL void load8_split(const __half* p, int lo_off, float (&f)[8]) {
  const uint4 uh = __ldg(reinterpret_cast<const uint4*>(p));
  const uint4 ul = __ldg(reinterpret_cast<const uint4*>(p + lo_off));
  float a[8], b[8];
  unpack8h(uh, a);
  unpack8h(ul, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = a[i] + b[i];
}
DEVINL void store8_split(__half* p, int lo_off, const float (&f)[8]) {
  __align__(16) __half hi[8];
  __align__(16) __half lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_h(f[i], hi[i], lo[i]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(p + lo_off) = *reinterpret_cast<const uint4*>(lo);
}

// ---------------------------------------------------------------------------------------------------------
// GEMM / implicit-GEMM conv with split-fp16 output.  Same tcgen05 main loop as gemm_conv_kernel (GemmArgs.fp16 = 1
// selects the half-precision instruction descriptor), plus SPLIT-K: contender batches are tiny (M = a few hundred rows at
// the 8x8 / 16x16 levels against K up to 41 472), so the K range of every output tile is cut into `splits` slices that run
// on different SMs; each slice stores its fp32 partial tile and gemm_prec_finish_kernel adds the slices IN ORDER and applies
// (acc * acc_scale + bias + residual) * out_scale -> hi / lo halves (or fp32).  `splits` is a function of the layer alone
// (never of the batch), so a sample's bits do not depend on how many contenders share its launch.  splits == 1: the
// epilogue finishes the tile itself.
// ---------------------------------------------------------------------------------------------------------
struct GemmPrecArgs {
  float acc_scale;               // weights are stored pre-multiplied by a power of two (keeps Wlo out of the half subnormals)
  int out_lo_off;                // column offset of the lo plane in `out` (halves); unused for fp32 output
  const __half* res;             // split-fp16 residual [M, ld_res] or null
  int ld_res, res_lo_off;
  int splits;                    // K slices per output tile
  int kb_per_split;              // K blocks (of 64) per slice
  float* partial;                // [splits][m_tiles*128][n_tiles*BN] fp32 (splits > 1)
  int ld_partial;                // n_tiles*BN
  int n_pad;                     // padded output columns (row pitch of the K-block-major weight matrix)
  int lo_off[3];                 // per activation source: channel offset of its lo plane (= its logical channel count)
  int* ticket;                   // reserved (an in-kernel finish by the last-arriving slice was measured 2.7x SLOWER than the
                                 // separate finishing launch: one row per thread is a serial chain of L2 round trips)
};

template <int BN>
struct GemmPrecCfg {
  using Base = GemmCfg<BN>;
  // One pipeline stage = one LOGICAL K block of 64 channels: the hi and the lo plane of the activation tile and the hi and
  // the lo plane of the weight tile, consumed by three MMAs per K slice (hi x Whi, lo x Whi, hi x Wlo).  Round 2's first
  // version ran the same three products as three K segments of the 16-bit main loop, i.e. fetched the hi activation tile
  // and the Whi tile TWICE (120 KB of operands per logical K block at BN = 192 against 80 KB here); small-row launches are
  // bound by exactly that L2 -> shared-memory stream.  No epilogue slot ring here, so the whole shared memory is the ring.
  static constexpr int A2_BYTES = 2 * Base::A_BYTES;
  static constexpr int B2_BYTES = 2 * Base::B_BYTES;
  static constexpr int STAGE_BYTES = A2_BYTES + B2_BYTES;
  static constexpr int STAGES = (BN >= 192) ? 2 : (BN >= 128 ? 3 : (BN >= 64 ? 4 : 5));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int THREADS = 192;
};

// work item `it` of this CTA -> (m tile, n tile, K slice); K slice fastest: the slices of one tile run side by side
DEVINL bool gemm_prec_item(const GemmArgs& a, const GemmPrecArgs& pa, int it, int& mt, int& nt, int& ks) {
  const int total = a.m_tiles * a.n_tiles * pa.splits;
  const int w = blockIdx.x + it * gridDim.x;
  if (w >= total) return false;
  ks = w % pa.splits;
  const int tl = w / pa.splits;
  mt = tl / a.n_tiles;
  nt = tl % a.n_tiles;
  return true;
}

template <int BN>
DEVINL void gemm_prec_producer(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                               const GemmArgs& a, const GemmPrecArgs& pa, uint8_t* smem_a, uint8_t* smem_b, uint64_t* full_bar,
                               uint64_t* empty_bar) {
  using Cfg = GemmCfg<BN>;
  using PCfg = GemmPrecCfg<BN>;
  constexpr int STAGES = PCfg::STAGES;
  int stage = 0;
  uint32_t phase = 0;
  int mt, nt, ks;
  for (int it = 0; gemm_prec_item(a, pa, it, mt, nt, ks); ++it) {
    const int kb0 = ks * pa.kb_per_split, kb1 = min(a.nkb, kb0 + pa.kb_per_split);
    int n0, y0, x0 = 0;
    if (a.tiles_per_img > 0) {
      n0 = mt / a.tiles_per_img;
      const int r = mt % a.tiles_per_img;
      y0 = (r / a.x_chunks) * a.tileH;
      x0 = (r % a.x_chunks) * 128;
    } else {
      n0 = mt * a.tileN;
      y0 = 0;
    }
    int kb = 0;
    for (int s = 0; s < a.n_seg; ++s) {
      const KSeg sg = a.seg[s];
      const CUtensorMap* tm = sg.src == 0 ? &tmA0 : (sg.src == 1 ? &tmA1 : &tmA2);
      if (kb + sg.taps * sg.cblocks <= kb0 || kb >= kb1) {      // the whole segment lies outside this slice
        kb += sg.taps * sg.cblocks;
        continue;
      }
      for (int tap = 0; tap < sg.taps; ++tap) {
        const int dy = sg.taps == 9 ? tap / 3 - 1 : 0;
        const int dx = sg.taps == 9 ? tap % 3 - 1 : 0;
        for (int cb = 0; cb < sg.cblocks; ++cb, ++kb) {
          if (kb < kb0 || kb >= kb1) continue;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], PCfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * PCfg::A2_BYTES;
          uint8_t* sb = smem_b + stage * PCfg::B2_BYTES;
          const int c_hi = sg.cstart + cb * 64;
          tma_load_4d(sa, tm, &full_bar[stage], c_hi, x0 + dx, y0 + dy, n0);
          tma_load_4d(sa + Cfg::A_BYTES, tm, &full_bar[stage], c_hi + pa.lo_off[sg.src], x0 + dx, y0 + dy, n0);
          // weights are stored K-block-major ([nkb][hi, lo][Npad][64]): the BN x 64 tile of one plane of K block kb is ONE
          // contiguous BN*128-byte run of HBM (row-major [Npad][Ktot] would be BN scattered 128-byte reads per tile,
          // which is what small-M launches -- they stream every weight exactly once -- are bound by)
          tma_load_2d(sb, &tmB, &full_bar[stage], 0, (2 * kb) * pa.n_pad + nt * BN);
          tma_load_2d(sb + Cfg::B_BYTES, &tmB, &full_bar[stage], 0, (2 * kb + 1) * pa.n_pad + nt * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  }
}

template <int BN>
DEVINL void gemm_prec_mma(const GemmArgs& a, const GemmPrecArgs& pa, uint8_t* smem_a, uint8_t* smem_b, uint64_t* full_bar,
                          uint64_t* empty_bar, uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t tmem_base) {
  using Cfg = GemmCfg<BN>;
  using PCfg = GemmPrecCfg<BN>;
  constexpr int STAGES = PCfg::STAGES;
  const uint32_t idesc = umma_idesc_f16(128, BN);
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  int mt, nt, ks;
  for (int it = 0; gemm_prec_item(a, pa, it, mt, nt, ks); ++it) {
    const int kb0 = ks * pa.kb_per_split, kb1 = min(a.nkb, kb0 + pa.kb_per_split);
    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc * BN;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint64_t da_hi = umma_desc_sw128(smem_u32(smem_a + stage * PCfg::A2_BYTES));
      const uint64_t da_lo = umma_desc_sw128(smem_u32(smem_a + stage * PCfg::A2_BYTES + Cfg::A_BYTES));
      const uint64_t db_hi = umma_desc_sw128(smem_u32(smem_b + stage * PCfg::B2_BYTES));
      const uint64_t db_lo = umma_desc_sw128(smem_u32(smem_b + stage * PCfg::B2_BYTES + Cfg::B_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) {       // (hi + lo) x (Whi + Wlo) without the 2^-22 lo x Wlo term, fp32 accumulation in TMEM
        umma_bf16(d_tmem, da_hi + 2 * k, db_hi + 2 * k, idesc, (kb > kb0) || (k != 0));
        umma_bf16(d_tmem, da_lo + 2 * k, db_hi + 2 * k, idesc, 1);
        umma_bf16(d_tmem, da_hi + 2 * k, db_lo + 2 * k, idesc, 1);
      }
      umma_commit(&empty_bar[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    umma_commit(&tfull_bar[acc]);
    if (++acc == 2) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
}

// (acc * acc_scale + bias + residual) * out_scale for 8 consecutive columns of row m -> split halves or fp32
DEVINL void gemm_prec_store8(const GemmArgs& a, const GemmPrecArgs& pa, int m, int n, float (&v)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    v[e] = __fmul_rn(v[e], pa.acc_scale);
    if (a.bias != nullptr) v[e] += __ldg(a.bias + n + e);
  }
  if (pa.res != nullptr) {
    float f[8];
    load8_split(pa.res + static_cast<size_t>(m) * pa.ld_res + n, pa.res_lo_off, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += f[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] *= a.out_scale;
  if (a.out_fp32) {
    float* o = reinterpret_cast<float*>(a.out) + static_cast<size_t>(m) * a.ld_out + n;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = v[e];
  } else {
    store8_split(reinterpret_cast<__half*>(a.out) + static_cast<size_t>(m) * a.ld_out + n, pa.out_lo_off, v);
  }
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_prec_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB, const GemmArgs a,
                 const GemmPrecArgs pa) {
  using Cfg = GemmCfg<BN>;
  using PCfg = GemmPrecCfg<BN>;
  constexpr int STAGES = PCfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * PCfg::A2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * PCfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                   // barriers, tensor maps and TMEM are set up while the predecessor drains

  if (warp == 0) {
    if (lane == 0) gemm_prec_producer<BN>(tmA0, tmA1, tmA2, tmB, a, pa, smem_a, smem_b, full_bar, empty_bar);
  } else if (warp == 1) {
    if (lane == 0) gemm_prec_mma<BN>(a, pa, smem_a, smem_b, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base);
  } else {
    // epilogue: warps 2..5, one per TMEM lane quarter (a warp may only touch lanes 32*(warp%4) ..)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int mt, nt, ks;
    for (int it = 0; gemm_prec_item(a, pa, it, mt, nt, ks); ++it) {
      const int m = mt * 128 + row;
      const bool m_ok = m < a.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if constexpr (BN == 16) {
        uint32_t r[16];
        tmem_ld16(t_row, r);
        tmem_ld_wait();
        if (pa.splits > 1) {
          float* o = pa.partial + (static_cast<size_t>(ks) * a.m_tiles * 128 + m) * pa.ld_partial + nt * BN;
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(r[j]);
        } else if (m_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = nt * BN + j;
            if (n < a.N) {
              float v = __fmul_rn(__uint_as_float(r[j]), pa.acc_scale);
              if (a.bias != nullptr) v += __ldg(a.bias + n);
              v *= a.out_scale;
              if (a.out_fp32) {
                reinterpret_cast<float*>(a.out)[static_cast<size_t>(m) * a.ld_out + n] = v;
              } else {
                __half hi, lo;
                split_h(v, hi, lo);
                __half* o = reinterpret_cast<__half*>(a.out) + static_cast<size_t>(m) * a.ld_out + n;
                o[0] = hi;
                o[pa.out_lo_off] = lo;
              }
            }
          }
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
          uint32_t r[32];
          tmem_ld32(t_row + j * 32, r);
          tmem_ld_wait();
          const int n0 = nt * BN + j * 32;
          if (pa.splits > 1) {                       // raw fp32 partial tile (rows past M are zero: TMA zero fill)
            float4* o = reinterpret_cast<float4*>(pa.partial + (static_cast<size_t>(ks) * a.m_tiles * 128 + m) * pa.ld_partial + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              o[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                 __uint_as_float(r[4 * i + 3]));
          } else if (m_ok && n0 < a.N) {             // N is a multiple of 32 for split output (checked on the host)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * i + e]);
              gemm_prec_store8(a, pa, m, n0 + 8 * i, v);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// splits > 1: out = finish(sum over slices in order).  One thread per (row, 8 columns); N % 8 == 0 or the fp32 N < 8 case.
__global__ void __launch_bounds__(256) gemm_prec_finish_kernel(const GemmArgs a, const GemmPrecArgs pa) {
  pdl_launch_dependents();      // programmatic dependent launch (small-row plans are latency-bound): the successor's prologue
  pdl_wait();                   // overlaps this kernel; nothing is read before the predecessor has completed
  const int n8 = (a.N + 7) >> 3;
  const long long total = static_cast<long long>(a.M) * n8;
  const size_t slice = static_cast<size_t>(a.m_tiles) * 128 * pa.ld_partial;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * 256) {
    const int m = static_cast<int>(idx / n8);
    const int n = static_cast<int>(idx % n8) * 8;
    const float* p = pa.partial + static_cast<size_t>(m) * pa.ld_partial + n;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < pa.splits; ++s) {
      const float4 x0 = *reinterpret_cast<const float4*>(p + s * slice);
      const float4 x1 = *reinterpret_cast<const float4*>(p + s * slice + 4);
      v[0] += x0.x; v[1] += x0.y; v[2] += x0.z; v[3] += x0.w;
      v[4] += x1.x; v[5] += x1.y; v[6] += x1.z; v[7] += x1.w;
    }
    if (n + 8 <= a.N) {
      gemm_prec_store8(a, pa, m, n, v);
    } else {                                         // ragged tail (the 3-channel fp32 output conv)
      for (int e = 0; n + e < a.N; ++e) {
        float y = __fmul_rn(v[e], pa.acc_scale);
        if (a.bias != nullptr) y += __ldg(a.bias + n + e);
        y *= a.out_scale;
        if (a.out_fp32) {
          reinterpret_cast<float*>(a.out)[static_cast<size_t>(m) * a.ld_out + n + e] = y;
        } else {
          __half hi, lo;
          split_h(y, hi, lo);
          __half* o = reinterpret_cast<__half*>(a.out) + static_cast<size_t>(m) * a.ld_out + n + e;
          o[0] = hi;
          o[pa.out_lo_off] = lo;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm over split-fp16 NHWC tensors (up to two concatenated sources): statistics in fp64, apply in fp32.
// ---------------------------------------------------------------------------------------------------------
struct GnPrecArgs {
  const __half* x0;
  const __half* x1;
  int C0, C1, C;             // logical channels (each source tensor is [.., 2*Ci] halves, lo plane at +Ci)
  int H, W, groups, cpg;     // INPUT spatial dims
  float eps;
  const float* gamma;
  const float* beta;
  const float* pre_add;      // [b_emb, ld_pre_add] added before the norm (networks.py:175), or null
  int ld_pre_add;
  const float* film_scale;   // [b_emb, ld_film] or null:  y = shift + norm * (scale + 1)   (networks.py:173)
  const float* film_shift;
  int ld_film, b_emb;
  int silu, resample;        // resample: 0 none, 1 = nearest 2x up, 2 = 2x2 mean down
  __half* out;               // split [batch, H', W', 2C]
  __half* raw_out;           // split, the resampled un-normalised input, or null
  float2* mean_rstd;         // [batch, groups]
  int batch;
  double* partial;           // [batch, splits, groups, 2] fp64 scratch of the statistics kernel
  int splits, PY;            // pixel splits per sample (a function of H*W only); pixel rows per CTA pass
  int* ticket;               // [batch] zero-initialised arrival counters (reset by the last CTA of a sample)
};

// grid (splits, batch), (C/8)*PY threads: fp64 sums of (hi + lo [+ pre_add]) per (sample, split, group) in a fixed order; the
// LAST CTA of a sample to finish (ticket) adds the splits in order and writes (mean, rstd) -- which CTA is last does not
// matter, the summation order is fixed, so the result is bit-identical at any batch position.
__global__ void __launch_bounds__(256) gn_stats_prec_kernel(const GnPrecArgs a) {
  __shared__ double s_sum[2048];
  __shared__ double s_sq[2048];
  __shared__ int s_last;
  pdl_launch_dependents();      // programmatic dependent launch (small-row plans are latency-bound): the successor's prologue
  pdl_wait();                   // overlaps this kernel; nothing is read before the predecessor has completed
  const int VC = a.C >> 3;
  const int vx = threadIdx.x % VC, py = threadIdx.x / VC;
  const int split = blockIdx.x, bi = blockIdx.y;
  const int HW = a.H * a.W;
  const int ppb = HW / a.splits;
  const int c = vx * 8;
  const __half* src;
  int Cs;
  if (c < a.C0) {
    src = a.x0 + c;
    Cs = a.C0;
  } else {
    src = a.x1 + (c - a.C0);
    Cs = a.C1;
  }
  src += (static_cast<size_t>(bi) * HW + static_cast<size_t>(split) * ppb) * (2 * Cs);
  float pa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (a.pre_add != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) pa[j] = a.pre_add[static_cast<size_t>(bi % a.b_emb) * a.ld_pre_add + c + j];
  }
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ss[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // 4 pixels (8 independent 16-byte loads) in flight per thread; the accumulation order stays p = py, py+PY, ...
  int p = py;
  for (; p + 3 * a.PY < ppb; p += 4 * a.PY) {
    uint4 uh[4], ul[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const __half* q = src + static_cast<size_t>(p + u * a.PY) * (2 * Cs);
      uh[u] = __ldg(reinterpret_cast<const uint4*>(q));
      ul[u] = __ldg(reinterpret_cast<const uint4*>(q + Cs));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float fh[8], fl[8];
      unpack8h(uh[u], fh);
      unpack8h(ul[u], fl);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double v = static_cast<double>((fh[j] + fl[j]) + pa[j]);
        s[j] += v;
        ss[j] += v * v;
      }
    }
  }
  for (; p < ppb; p += a.PY) {
    float f[8];
    load8_split(src + static_cast<size_t>(p) * (2 * Cs), Cs, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double v = static_cast<double>(f[j] + pa[j]);
      s[j] += v;
      ss[j] += v * v;
    }
  }
  // level 1: per-channel sums over py (fixed order) through shared memory, PY*C <= 2048 doubles
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_sum[py * a.C + c + j] = s[j];
    s_sq[py * a.C + c + j] = ss[j];
  }
  __syncthreads();
  // level 2: thread g adds its group's channels (and pixel rows) in a fixed order
  for (int g = threadIdx.x; g < a.groups; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int cc = g * a.cpg; cc < (g + 1) * a.cpg; ++cc)
      for (int y = 0; y < a.PY; ++y) {
        ds += s_sum[y * a.C + cc];
        dq += s_sq[y * a.C + cc];
      }
    double* o = a.partial + ((static_cast<size_t>(bi) * a.splits + split) * a.groups + g) * 2;
    o[0] = ds;
    o[1] = dq;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(a.ticket + bi, 1) == a.splits - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int g = threadIdx.x; g < a.groups; g += blockDim.x) {
    double ts = 0.0, tq = 0.0;
    const double2* o = reinterpret_cast<const double2*>(a.partial) + static_cast<size_t>(bi) * a.splits * a.groups + g;
    for (int sp0 = 0; sp0 < a.splits; sp0 += 8) {          // 8 independent L2 loads in flight, added in split order
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = sp0 + u < a.splits ? __ldcg(o + static_cast<size_t>(sp0 + u) * a.groups) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        ts += v[u].x;
        tq += v[u].y;
      }
    }
    const double n = static_cast<double>(HW) * a.cpg;
    const double mean = ts / n;
    double var = tq / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    a.mean_rstd[static_cast<size_t>(bi) * a.groups + g] =
        make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps))));
  }
  if (threadIdx.x == 0) a.ticket[bi] = 0;
}

DEVINL float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// one thread per (output pixel, 8-channel vector)
__global__ void __launch_bounds__(256) gn_apply_prec_kernel(const GnPrecArgs a) {
  pdl_launch_dependents();      // programmatic dependent launch (small-row plans are latency-bound): the successor's prologue
  pdl_wait();                   // overlaps this kernel; nothing is read before the predecessor has completed
  const int VC = a.C >> 3;
  const int outW = a.resample == 1 ? a.W * 2 : (a.resample == 2 ? a.W / 2 : a.W);
  const int outH = a.resample == 1 ? a.H * 2 : (a.resample == 2 ? a.H / 2 : a.H);
  const long long total = static_cast<long long>(a.batch) * outH * outW * VC;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * 256) {
    const int vx = static_cast<int>(idx % VC);
    long long t = idx / VC;
    const int ox = static_cast<int>(t % outW);
    t /= outW;
    const int oy = static_cast<int>(t % outH);
    const int bi = static_cast<int>(t / outH);
    const int c = vx * 8;
    const __half* src;
    int Cs;
    if (c < a.C0) {
      src = a.x0 + c;
      Cs = a.C0;
    } else {
      src = a.x1 + (c - a.C0);
      Cs = a.C1;
    }
    src += static_cast<size_t>(bi) * a.H * a.W * (2 * Cs);
    float ka[8], kb[8], pad[8];
    const int e = bi % a.b_emb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cc = c + j;
      const float2 mr = a.mean_rstd[static_cast<size_t>(bi) * a.groups + cc / a.cpg];
      const float rs = mr.y * a.gamma[cc];
      float k1 = rs;
      float k0 = a.beta[cc] - mr.x * rs;
      pad[j] = a.pre_add != nullptr ? a.pre_add[static_cast<size_t>(e) * a.ld_pre_add + cc] : 0.f;
      if (a.film_scale != nullptr) {
        const float sc = a.film_scale[static_cast<size_t>(e) * a.ld_film + cc] + 1.0f;
        const float sh = a.film_shift[static_cast<size_t>(e) * a.ld_film + cc];
        k1 *= sc;
        k0 = k0 * sc + sh;
      }
      ka[j] = k1;
      kb[j] = k0;
    }
    float v[8], raw[8];
    if (a.resample == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = raw[j] = 0.f;
#pragma unroll
      for (int tq = 0; tq < 4; ++tq) {
        float f[8];
        load8_split(src + static_cast<size_t>((2 * oy + (tq >> 1)) * a.W + 2 * ox + (tq & 1)) * (2 * Cs), Cs, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = (f[j] + pad[j]) * ka[j] + kb[j];
          if (a.silu) y = silu_exact(y);
          v[j] += y;
          raw[j] += f[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] *= 0.25f;
        raw[j] *= 0.25f;
      }
    } else {
      const int iy = a.resample == 1 ? oy >> 1 : oy, ix = a.resample == 1 ? ox >> 1 : ox;
      load8_split(src + static_cast<size_t>(iy * a.W + ix) * (2 * Cs), Cs, raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = (raw[j] + pad[j]) * ka[j] + kb[j];
        if (a.silu) y = silu_exact(y);
        v[j] = y;
      }
    }
    const size_t o = (static_cast<size_t>(bi) * outH * outW + static_cast<size_t>(oy) * outW + ox) * (2 * a.C) + c;
    store8_split(a.out + o, a.C, v);
    if (a.raw_out != nullptr) store8_split(a.raw_out + o, a.C, raw);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Self-attention, head_dim 64, fp32 on the FMA pipe (AttentionOp, networks.py:113-118: fp32 softmax of q . k/sqrt(64)).
// One CTA = 64 queries of one (sample, head); K/V stream through shared memory in tiles of 64 keys; online softmax
// with exact expf; thread (ty, tx) owns the 4x4 block [queries 4ty.., keys/dims 4tx..].
// ---------------------------------------------------------------------------------------------------------
struct AttnPrecArgs {
  const __half* qkv;         // split [batch*L, ld] : hi plane columns [0, lo_off), lo plane at +lo_off
  int ld, lo_off;
  int k_col0, v_col0;        // Q at column head*64, K at k_col0 + head*64, V at v_col0 + head*64 (inside a plane)
  __half* out;               // split [batch*L, ld_out], head h at column h*64
  int ld_out, out_lo_off;
  int heads, L;
  float scale;
};

constexpr int ATTN_PREC_PITCH = 68;
constexpr int ATTN_PREC_SMEM = (3 * 64 * ATTN_PREC_PITCH + 64 * 64) * 4;

__global__ void __launch_bounds__(256) attention_prec_kernel(const AttnPrecArgs a) {
  extern __shared__ float sm[];
  float* Qs = sm;                                  // [d][q]   pitch 68
  float* Ks = Qs + 64 * ATTN_PREC_PITCH;           // [d][key] pitch 68
  float* Ps = Ks + 64 * ATTN_PREC_PITCH;           // [key][q] pitch 68
  float* Vs = Ps + 64 * ATTN_PREC_PITCH;           // [key][d] pitch 64
  pdl_launch_dependents();      // programmatic dependent launch (small-row plans are latency-bound): the successor's prologue
  pdl_wait();                   // overlaps this kernel; nothing is read before the predecessor has completed
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int bh = blockIdx.y;
  const int bi = bh / a.heads, h = bh - bi * a.heads;
  const int q0 = blockIdx.x * 64;
  const size_t row0 = static_cast<size_t>(bi) * a.L;

  // Q tile -> Qs[d][q] (pre-scaled: the reference scales k by 1/sqrt(64), an exact power of two)
  // (consecutive lanes take consecutive rows: the transposed scalar stores are bank-conflict free)
  for (int v = tid; v < 64 * 8; v += 256) {
    const int r = v & 63, d8 = (v >> 6) * 8;
    float f[8];
    load8_split(a.qkv + (row0 + q0 + r) * a.ld + h * 64 + d8, a.lo_off, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) Qs[(d8 + j) * ATTN_PREC_PITCH + r] = f[j] * a.scale;
  }
  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < a.L; k0 += 64) {
    __syncthreads();                               // previous tile's Ks / Ps / Vs fully consumed (and Qs written)
    for (int v = tid; v < 64 * 8; v += 256) {
      const int r = v & 63, d8 = (v >> 6) * 8;
      float f[8];
      load8_split(a.qkv + (row0 + k0 + r) * a.ld + a.k_col0 + h * 64 + d8, a.lo_off, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) Ks[(d8 + j) * ATTN_PREC_PITCH + r] = f[j];
    }
    for (int v = tid; v < 64 * 8; v += 256) {      // V keeps its [key][d] layout: two float4 stores per thread
      const int r = v >> 3, d8 = (v & 7) * 8;
      float f[8];
      load8_split(a.qkv + (row0 + k0 + r) * a.ld + a.v_col0 + h * 64 + d8, a.lo_off, f);
      *reinterpret_cast<float4*>(Vs + r * 64 + d8) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(Vs + r * 64 + d8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      const float4 qv = *reinterpret_cast<const float4*>(Qs + d * ATTN_PREC_PITCH + 4 * ty);
      const float4 kv = *reinterpret_cast<const float4*>(Ks + d * ATTN_PREC_PITCH + 4 * tx);
      const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
    }
    // online softmax: the 16 threads of a row group (same ty) are 16 consecutive lanes
    float corr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);
      corr[i] = expf(m_run[i] - m_new);            // first tile: exp(-inf) = 0
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = expf(s[i][j] - m_new);
        rs += s[i][j];
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[i] = l_run[i] * corr[i] + rs;
      m_run[i] = m_new;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(Ps + (4 * tx + j) * ATTN_PREC_PITCH + 4 * ty) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr[i];
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 64; ++kk) {
      const float4 pv = *reinterpret_cast<const float4*>(Ps + kk * ATTN_PREC_PITCH + 4 * ty);
      const float4 vv = *reinterpret_cast<const float4*>(Vs + kk * 64 + 4 * tx);
      const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pa[i], va[j], o[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half* op = a.out + (row0 + q0 + 4 * ty + i) * a.ld_out + h * 64 + 4 * tx;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __half hi, lo;
      split_h(o[i][j] / l_run[i], hi, lo);
      op[j] = hi;
      op[a.out_lo_off + j] = lo;
    }
  }
}

// 3x3 im2col of the fp32 NCHW network input (Cin*9 <= 64) into split [batch*H*W, 128]: hi taps | lo taps
__global__ void im2col_c3_prec_kernel(const float* __restrict__ x, __half* __restrict__ out, int batch, int C, int H, int W) {
  pdl_launch_dependents();      // programmatic dependent launch (small-row plans are latency-bound): the successor's prologue
  pdl_wait();                   // overlaps this kernel; nothing is read before the predecessor has completed
  const long long total = static_cast<long long>(batch) * H * W * 64;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * 256) {
    const int k = static_cast<int>(idx & 63);
    long long p = idx >> 6;
    const int xw = static_cast<int>(p % W);
    p /= W;
    const int yh = static_cast<int>(p % H);
    const int bi = static_cast<int>(p / H);
    float v = 0.f;
    if (k < 9 * C) {
      const int tap = k / C, c = k - tap * C;
      const int yy = yh + tap / 3 - 1, xx = xw + tap % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((static_cast<size_t>(bi) * C + c) * H + yy) * W + xx];
    }
    __half hi, lo;
    split_h(v, hi, lo);
    __half* o = out + (idx >> 6) * 128 + k;
    o[0] = hi;
    o[64] = lo;
  }
}

}  // namespace b200
