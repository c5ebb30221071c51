// Implicit-GEMM convolution (3x3 pad 1 / 1x1) and plain GEMM on tcgen05 tensor cores.
//
//   out[m, n] = out_scale * ( sum_k A[m,k] * W[n,k] + bias[n] + residual[m,n] )
//
// A is never materialised: for every K block (one filter tap x 64 input channels) the TMA
// producer loads a [tileN, tileH, tileW, 64ch] box of the NHWC activation tensor, shifted
// by the tap offset, straight into a 128-row x 128-byte SWIZZLE_128B shared-memory tile
// (out-of-bounds pixels are zero-filled by TMA == the conv's zero padding).  Several K
// segments can be chained (channel concat of two tensors; conv1 + 1x1 skip conv fused into
// one accumulator).  Accumulators live in TMEM (double buffered), so the epilogue of tile i
// overlaps the main loop of tile i+1.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..9 = epilogue (TMEM -> registers -> global): two warps per TMEM lane quarter, each taking
// half of the tile's columns, residual rows prefetched one chunk ahead.
// Replaces Conv2d.forward, edm/training/networks.py:68-90 (kernel 3 / 1, no resample).
#pragma once
#include "common.cuh"

namespace b200 {

struct KSeg {
  int src, taps, cstart, cblocks;
};

struct GemmArgs {
  int n_seg;
  KSeg seg[4];
  int nkb;                       // total K blocks of 64
  int M, N;                      // valid rows / cols
  int H, W;                      // spatial dims
  int tileH, tileN;              // box = [tileN][tileH][W] pixels = 128 rows
  int tiles_per_img;             // H*W/128 (0 if an image is smaller than a tile)
  int m_tiles, n_tiles;
  const float* bias;
  const __nv_bfloat16* residual;
  int ld_res;
  float out_scale;
  void* out;
  int ld_out;
  int out_fp32;
  __nv_bfloat16* vt_out;
  int vt_col_start, heads, L;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;      // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 192 ? 5 : 6);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int STAGING_BYTES = 8 * 2048;   // per-epilogue-warp 32 rows x 64 B transpose buffer
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int THREADS = 320;
};

template <int BN>
__global__ void __launch_bounds__(320, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB, const GemmArgs a) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_stg = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = a.m_tiles * a.n_tiles;

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
      mbar_init(&tempty_bar[i], 8);
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / a.n_tiles, nt = tile % a.n_tiles;
        int n0, y0;
        if (a.tiles_per_img > 0) {
          n0 = mt / a.tiles_per_img;
          y0 = (mt % a.tiles_per_img) * a.tileH;
        } else {
          n0 = mt * a.tileN;
          y0 = 0;
        }
        int kb = 0;
        for (int s = 0; s < a.n_seg; ++s) {
          const KSeg sg = a.seg[s];
          const CUtensorMap* tm = sg.src == 0 ? &tmA0 : (sg.src == 1 ? &tmA1 : &tmA2);
          for (int tap = 0; tap < sg.taps; ++tap) {
            const int dy = sg.taps == 9 ? tap / 3 - 1 : 0;
            const int dx = sg.taps == 9 ? tap % 3 - 1 : 0;
            for (int cb = 0; cb < sg.cblocks; ++cb, ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              tma_load_4d(smem_a + stage * Cfg::A_BYTES, tm, &full_bar[stage], sg.cstart + cb * 64, dx, y0 + dy, n0);
              tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * 64, nt * BN);
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < a.nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * Cfg::A_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (addr>>4) field
            umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
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
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which half of the tile's columns
    const int row = q * 32 + lane;          // row inside the 128-row tile
    constexpr int CH = (BN % 32 == 0) ? 32 : 16;
    constexpr int NCH = BN / CH;
    constexpr int CH_PER_HALF = (NCH + 1) / 2;
    const int ch_begin = half * CH_PER_HALF;
    const int ch_end = (ch_begin + CH_PER_HALF < NCH) ? ch_begin + CH_PER_HALF : NCH;
    const bool has_res = a.residual != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / a.n_tiles, nt = tile % a.n_tiles;
      const int m = mt * 128 + row;
      const bool m_ok = m < a.M;
      uint4 rnext[CH / 8];
      auto load_res = [&](int ch, uint4 (&dst)[CH / 8]) {
        const int nb = nt * BN + ch * CH;
        if (has_res && m_ok && nb + CH <= a.N) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.residual + static_cast<size_t>(m) * a.ld_res + nb);
#pragma unroll
          for (int j = 0; j < CH / 8; ++j) dst[j] = __ldg(rp + j);
        }
      };
      if (ch_begin < ch_end) load_res(ch_begin, rnext);     // in flight while the main loop still runs
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        const int c0 = ch * CH;
        uint4 rcur[CH / 8];
#pragma unroll
        for (int j = 0; j < CH / 8; ++j) rcur[j] = rnext[j];
        if (ch + 1 < ch_end) load_res(ch + 1, rnext);
        float v[CH];
        if constexpr (CH == 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        } else {
          uint32_t r[16];
          tmem_ld16(t_row + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        }
        const int n_base = nt * BN + c0;
        if (n_base >= a.N) continue;                       // warp-uniform
        const bool full = (n_base + CH <= a.N);
        if (a.bias != nullptr) {
          if (full) {      // warp-uniform address: 16-byte broadcast loads
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n_base) + j);
              v[4 * j + 0] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (n_base + j < a.N) v[j] += __ldg(a.bias + n_base + j);
          }
        }
        if (has_res && full) {
#pragma unroll
          for (int j = 0; j < CH / 8; ++j) {
            const uint4 u = rcur[j];
            float2 f;
            f = unpack_bf16(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
            f = unpack_bf16(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
            f = unpack_bf16(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
            f = unpack_bf16(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
          }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] *= a.out_scale;

        if (a.out_fp32) {
          if (m_ok) {
            float* op = reinterpret_cast<float*>(a.out) + static_cast<size_t>(m) * a.ld_out + n_base;
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (full || n_base + j < a.N) op[j] = v[j];
          }
        } else if (a.vt_out != nullptr && n_base >= a.vt_col_start) {
          if (!m_ok) continue;
          // V^T[(batch*heads + head), d, p] : consecutive lanes -> consecutive pixels p
          const int vc = n_base - a.vt_col_start;
          const int head = vc >> 6, d0 = vc & 63;
          const int bi = m / a.L, p = m - bi * a.L;
          __nv_bfloat16* vp = a.vt_out + (static_cast<size_t>(bi * a.heads + head) * 64 + d0) * a.L + p;
#pragma unroll
          for (int j = 0; j < CH; ++j) vp[static_cast<size_t>(j) * a.L] = __float2bfloat16(v[j]);
        } else if (full && CH == 32) {
          // transpose through a per-warp smem buffer so that 4 consecutive lanes write one row's 64 B:
          // every global store instruction covers 8 rows x 64 contiguous bytes (full 32 B sectors)
          uint8_t* stg = smem_stg + (warp - 2) * 2048;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
            u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
            u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
            u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ sw) << 4)) = u;
          }
          __syncwarp();
          __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(a.out) + n_base + (lane & 3) * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (lane >> 2);
            const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4));
            const int mrow = mt * 128 + q * 32 + rr;
            if (mrow < a.M) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(mrow) * a.ld_out) = val;
          }
          __syncwarp();
        } else if (m_ok) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(m) * a.ld_out + n_base;
          for (int j = 0; j < CH; ++j)
            if (n_base + j < a.N) op[j] = __float2bfloat16(v[j]);
        }
      }
      // release this accumulator back to the MMA warp
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

}  // namespace b200
