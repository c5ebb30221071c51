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
// warps 2..9 = epilogue.
//
// gemm_conv_kernel<BN> (BN 64..256, bf16 output) -- the epilogue works on 128-row x 64-column
// sub-boxes through a 4-slot ring of 16 KB SWIZZLE_128B shared-memory tiles:
//   * the residual sub-box is TMA-loaded into the slot two sub-boxes ahead (no per-thread strided loads),
//   * TMEM -> registers (+bias, +residual from the slot, *scale) -> bf16 written back into the slot,
//   * one elected thread TMA-stores the slot (full 128-byte lines, asynchronous bulk group),
//   * optionally the per-channel sum / sum of squares of the STORED bf16 values over each 64-row half
//     tile are reduced from the slot in a fixed order and written to gn_stats[M/64, N] (float2): the
//     consumer's GroupNorm statistics come for free instead of costing another pass over HBM.
// gemm_small_n_kernel (BN = 16; fp32 or bf16 output, direct global stores) serves the 3-channel output convs.
// Replaces Conv2d.forward, edm/training/networks.py:68-90 (kernel 3 / 1, no resample).
#pragma once
#include "common.cuh"

namespace b200 {

struct KSeg {
  int src, taps, cstart, cblocks;
  int dy0, dx0;                  // taps == 4 (one phase of conv3x3 o nearest-upsample-2x): tap t reads pixel (y + dy0 + t/2, x + dx0 + t%2)
};

struct GemmArgs {
  int n_seg;
  KSeg seg[8];
  int nkb;                       // total K blocks of 64
  int M, N;                      // valid rows / cols
  int H, W;                      // spatial dims
  int tileH, tileN;              // box = [tileN][tileH][W] pixels = 128 rows
  int tiles_per_img;             // H*W/128 (0 if an image is smaller than a tile)
  int x_chunks;                  // W/128 when a row is wider than a tile (VAE decoder, W = 256 / 512): tile = 128 pixels of ONE row
  int m_tiles, n_tiles;
  const float* bias;
  const act_t* residual;
  int ld_res;
  float out_scale;
  void* out;
  int ld_out;
  int out_fp32;
  float2* gn_stats;              // [M/64, ld_stats] (sum, sum of squares) per 64-row half tile, or null
  int ld_stats;                  // row pitch of gn_stats in channels (= the full N when this launch covers a column slice)
  int src_stride[3];             // 1, or 2: the source is sampled with stride 2 (3x3 stride-2 pad-1 conv: Downsample2D)
  int reverse;                   // walk the tiles last-to-first (start on what the producer of A wrote last: L2 hits)
  // Fused "nearest 2x upsample, then 3x3 conv" (Conv2d(up=True) networks.py:72-80, Upsample2D upsampling.py): the output
  // pixel (2y+py, 2x+px) only ever sees a 2x2 neighbourhood of the LOW-resolution input, with the 3x3 weights that fall on
  // the same source pixel pre-summed -- 4 launches (one per phase (py, px)) of a 2x2-tap conv over the low-res tensor:
  // 16 instead of 36 MACs per input channel and output pixel, and the upsampled tensor is never written.
  int out4d;                     // 1: `out` is stored through a 4-D map [N, W, H, batch] whose pixel strides skip every other
                                 //    row / column of the high-res tensor (base shifted by the phase)
  int stats_in_rows;             // gn_stats rows (64 pixels each) per image in THIS launch (0: plain layout)
  int stats_img_rows;            // gn_stats rows per image in the high-res tensor (4 x stats_in_rows)
  int stats_off;                 // first gn_stats row of this phase inside an image
  int geglu;                     // 1: the weight rows come in groups of 128 = [64 hidden | 64 gate] and the epilogue stores
                                 //    hidden * gelu(gate) (exact erf GELU) as bf16 [M, N/2]: GEGLU (activations.py:117-123)
                                 //    fused into the projection, whose [M, N] output never touches HBM
  int fp16;                      // 1: operands are IEEE half (the split-fp16 "precise" path, precise.cuh) instead of bf16
  int act;                       // 1: quick_gelu(acc + bias) = x * sigmoid(1.702 x) before the residual / out_scale (CLIP MLP)
  // XF kernels only: GroupNorm (affine, no activation) of the A operand, applied to the TMA-landed shared-memory stage in
  // place by four transform warps before the MMA warp consumes it -- the `norm -> 1x1 conv` pairs (attention norm2 -> qkv,
  // networks.py:182-183; SD / VAE `norm -> proj_in`) without the normalised tensor ever being written:
  //   a'[m, c] = a[m, c] * (rstd[s, g] * gamma[c]) + (beta[c] - mean[s, g] * rstd[s, g] * gamma[c]),  s = sample of row m
  const float2* xf_mean_rstd;    // [batch, xf_groups] (gn_finalize_kernel)
  const float* xf_gamma;
  const float* xf_beta;
  int xf_groups, xf_cpg;
  int xf_batch;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;      // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // measured on B200: 3/4 stages lose nothing against 4/5 for BN 256/192 (the loads come from L2)
  static constexpr int STAGES = (BN >= 256) ? 3 : (BN >= 192 ? 4 : (BN >= 128 ? 5 : 6));
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SLOTS = 4;
  static constexpr int SLOT_BYTES = 128 * 128;     // 128 rows x 64 bf16
  static constexpr int NSUB = BN / 64;
  static constexpr int EPI_BYTES = (BN >= 64) ? SLOTS * SLOT_BYTES + 2 * BN * 4 : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int THREADS = 320;
};

// Tile schedule.  CL2 = false: tile (mt, nt) = blockIdx.x + it * gridDim.x, nt fastest.  CL2 = true (clusters of two CTAs):
// the pair walks "super tiles" (two M-adjacent tiles of the SAME column slice), CTA rank r taking mt = 2*mtp + r, so that
// both CTAs need the same weight tile at the same time and each fetches half of it for both (multicast).
template <bool CL2>
DEVINL bool gemm_tile_at(const GemmArgs& a, int it, int& mt, int& nt) {
  if constexpr (CL2) {
    const int pairs = gridDim.x >> 1, pid = blockIdx.x >> 1;
    const int S = ((a.m_tiles + 1) >> 1) * a.n_tiles;
    const int st = pid + it * pairs;
    if (st >= S) return false;
    const int sl = a.reverse ? S - 1 - st : st;
    mt = 2 * (sl / a.n_tiles) + static_cast<int>(cluster_ctarank());
    nt = sl % a.n_tiles;
    return true;
  } else {
    const int num_tiles = a.m_tiles * a.n_tiles;
    const int tile = blockIdx.x + it * gridDim.x;
    if (tile >= num_tiles) return false;
    const int tl = a.reverse ? num_tiles - 1 - tile : tile;
    mt = tl / a.n_tiles;
    nt = tl % a.n_tiles;
    return true;
  }
}

// ===================== TMA producer (one thread) =====================
template <int BN, bool CL2 = false, bool CG2 = false, bool XF = false>
DEVINL void gemm_producer(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                          const GemmArgs& a, uint8_t* smem_a, uint8_t* smem_b, uint64_t* full_bar, uint64_t* empty_bar) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES - (XF ? 1 : 0);      // XF: the last stage's shared memory holds the coefficient tables
  const int rank = CL2 ? static_cast<int>(cluster_ctarank()) : 0;
  int stage = 0;
  uint32_t phase = 0;
  int mt, nt;
  for (int it = 0; gemm_tile_at<CL2>(a, it, mt, nt); ++it) {
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
      const int sdn = a.src_stride[sg.src];          // stride 2: input row = 2*out_row + dy (TMA elementStrides = 2)
      for (int tap = 0; tap < sg.taps; ++tap) {
        const int dy = sg.taps == 9 ? tap / 3 - 1 : (sg.taps == 4 ? sg.dy0 + (tap >> 1) : 0);
        const int dx = sg.taps == 9 ? tap % 3 - 1 : (sg.taps == 4 ? sg.dx0 + (tap & 1) : 0);
        for (int cb = 0; cb < sg.cblocks; ++cb, ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (CG2) {
            // CTA pair: my 128 rows of A and MY HALF of the weight tile go into my own shared memory; the bytes are
            // counted on the LEADER's full barrier, which therefore completes when both CTAs' operands have landed
            const uint32_t full0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
            mbar_arrive_expect_tx_cluster(full0, Cfg::A_BYTES + Cfg::B_BYTES / 2);
            tma_load_4d_cg2(smem_a + stage * Cfg::A_BYTES, tm, full0, sg.cstart + cb * 64, x0 * sdn + dx, y0 * sdn + dy, n0);
            tma_load_2d_cg2(smem_b + stage * Cfg::B_BYTES, &tmB, full0, kb * 64, nt * BN + rank * (BN / 2));
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_4d(smem_a + stage * Cfg::A_BYTES, tm, &full_bar[stage], sg.cstart + cb * 64, x0 * sdn + dx, y0 * sdn + dy, n0);
          if constexpr (CL2) {            // tmB's box is BN/2 rows: my half of the weight tile, for both CTAs of the pair
            tma_load_2d_mc(smem_b + stage * Cfg::B_BYTES + rank * (Cfg::B_BYTES / 2), &tmB, &full_bar[stage], kb * 64,
                           nt * BN + rank * (BN / 2), 3);
          } else {
            tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * 64, nt * BN);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  }
}

// ===================== MMA issuer (one thread) =====================
template <int BN, bool CL2 = false, bool CG2 = false, bool XF = false>
DEVINL void gemm_mma(const GemmArgs& a, uint8_t* smem_a, uint8_t* smem_b, uint64_t* full_bar, uint64_t* empty_bar,
                     uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t tmem_base) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES - (XF ? 1 : 0);
  const uint32_t idesc = CG2 ? umma_idesc_act(256, BN) : (a.fp16 ? umma_idesc_f16(128, BN) : umma_idesc_act(128, BN));
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  int mt_, nt_;
  for (int it = 0; gemm_tile_at<CL2>(a, it, mt_, nt_); ++it) {
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
        if constexpr (CG2)
          umma_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);     // M = 256: my rows + the peer's, B = both halves
        else
          umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
      }
      if constexpr (CG2)
        umma_commit_cg2_mc(&empty_bar[stage], 3);  // both CTAs' stages are free once the pair's MMAs have read them
      else if constexpr (CL2)
        umma_commit_mc(&empty_bar[stage], 3);      // the peer writes half of this stage too: release it in both CTAs
      else
        umma_commit(&empty_bar[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if constexpr (CG2)
      umma_commit_cg2_mc(&tfull_bar[acc], 3);      // each CTA's epilogue reads its own 128 accumulator rows
    else
      umma_commit(&tfull_bar[acc]);
    if (++acc == 2) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Main kernel: BN in {64,128,192,256}, bf16 output through the TMA slot ring.
// ---------------------------------------------------------------------------------------------------------
// CG2 (with CL2 scheduling): tcgen05 cta_group::2 -- the pair's leader issues M = 256 MMAs over both CTAs' A rows and the
// two halves of the weight tile, one half in each CTA's shared memory (half the B staging and operand traffic per SM).
// XF: four more warps (threads 320..447) normalise every A stage in place (GemmArgs.xf_*); the MMA warp then waits for
// THEIR barrier instead of the TMA's.  Single-source 1x1 GEMMs only, no clusters.
template <int BN>
DEVINL void gemm_transform(const GemmArgs& a, uint8_t* smem_a, float2* tab, uint64_t* full_bar, uint64_t* xf_bar) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES - 1;
  const int tt = threadIdx.x - 320;              // 0..127
  const int j = tt & 7;                          // 16-byte chunk (8 channels) of the 64-channel K block
  const int rbase = tt >> 3;                     // rows rbase + 16 i, i = 0..7
  const int HW = a.H * a.W;
  const int C = a.nkb * 64;                      // one 1x1 segment over all channels
  int stage = 0;
  uint32_t phase = 0;
  int mt, nt;
  for (int it = 0; gemm_tile_at<false>(a, it, mt, nt); ++it) {
    // samples of this tile: one (HW >= 128), or two of 64 rows each (HW = 64)
    const int s0 = a.tiles_per_img > 0 ? mt / a.tiles_per_img : mt * a.tileN;
    const int n_s = a.tiles_per_img > 0 ? 1 : 2;
    const int rows_per_sample = a.tiles_per_img > 0 ? 128 : HW;
    // (k_a, k_b) of every channel for the tile's samples -> table `it & 1` (double buffered: a thread that is already in
    // tile it + 1 writes the other table; the barrier below keeps everybody within one tile of each other)
    float2* t0 = tab + static_cast<size_t>(it & 1) * 2 * C;
    for (int c = tt; c < C; c += 128) {
      const float gam = __ldg(a.xf_gamma + c), bet = __ldg(a.xf_beta + c);
      const int g = c / a.xf_cpg;
      for (int u = 0; u < n_s; ++u) {
        int sidx = s0 + u;
        if (sidx >= a.xf_batch) sidx = a.xf_batch - 1;             // rows past M (zero-filled by the TMA, never stored)
        const float2 mr = __ldg(a.xf_mean_rstd + static_cast<size_t>(sidx) * a.xf_groups + g);
        const float rs = mr.y * gam;                               // same expressions as gn_apply_kernel: bit-identical
        // layout [K block][e = channel pair 0..3][j = 16-byte chunk 0..7][2]: the eight threads of a quarter warp (same e,
        // j = 0..7) read 128 contiguous bytes -- channel-major order was a 4-way bank conflict on every table read
        const int cc = c & 63;
        t0[u * C + (c & ~63) + ((cc & 7) >> 1) * 16 + (cc >> 3) * 2 + (cc & 1)] = make_float2(rs, bet - mr.x * rs);
      }
    }
    named_barrier_sync(2, 128);
    for (int kb = 0; kb < a.nkb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      const uint32_t base = smem_u32(smem_a + stage * Cfg::A_BYTES);
#pragma unroll
      for (int h = 0; h < 2; ++h) {              // rows 0..63 (sample 0) / 64..127 (sample 0 or 1)
        const int u = (h * 64 >= rows_per_sample) ? 1 : 0;
        const float4* tp = reinterpret_cast<const float4*>(t0 + u * C + kb * 64) + j;
        float4 k[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) k[e] = tp[e * 8];              // (ka, kb) of channels j*8 + 2e, j*8 + 2e + 1
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rbase + 16 * (4 * h + i);
          const uint32_t addr = base + static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(j ^ (r & 7)) << 4);
          uint4 v = lds128(addr);
          float2 f;
          f = unpack_act(v.x); v.x = pack_act(f.x * k[0].x + k[0].y, f.y * k[0].z + k[0].w);
          f = unpack_act(v.y); v.y = pack_act(f.x * k[1].x + k[1].y, f.y * k[1].z + k[1].w);
          f = unpack_act(v.z); v.z = pack_act(f.x * k[2].x + k[2].y, f.y * k[2].z + k[2].w);
          f = unpack_act(v.w); v.w = pack_act(f.x * k[3].x + k[3].y, f.y * k[3].z + k[3].w);
          sts128(addr, v);
        }
      }
      fence_proxy_async_smem();                  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if ((tt & 31) == 0) mbar_arrive(&xf_bar[stage]);   // one arrival per warp (128 arrivals on one barrier serialise)
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  }
}

template <int BN, bool CL2 = false, bool CG2 = false, bool XF = false>
__global__ void __launch_bounds__(XF ? 448 : 320, 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const GemmArgs a) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES - (XF ? 1 : 0);   // XF: one stage less, its shared memory = the coefficient tables
  constexpr int NSUB = Cfg::NSUB;
  constexpr int SLOTS = Cfg::SLOTS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  float2* xf_tab = reinterpret_cast<float2*>(smem + STAGES * Cfg::STAGE_BYTES);       // [2 tiles][2 samples][C] (k_a, k_b), XF only
  uint8_t* smem_slot = smem + Cfg::STAGES * Cfg::STAGE_BYTES;             // 1024-aligned (all sizes are multiples of 1 KB)
  float* smem_bias = reinterpret_cast<float*>(smem_slot + SLOTS * Cfg::SLOT_BYTES);   // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_slot + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint64_t* rfull_bar = bars + 2 * STAGES + 4;                            // [SLOTS] residual sub-box landed
  uint64_t* xf_bar = bars + 2 * STAGES + 4 + SLOTS;                       // [STAGES] A stage normalised in place (XF)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4 + SLOTS);
  static_assert(!(XF && CL2), "the transform warps are not cluster-aware");
  static_assert((3 * STAGES + 4 + SLOTS) * 8 + 4 <= 256, "barrier block");

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool has_res = a.residual != nullptr;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmO);
    if (has_res) prefetch_tmap(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], CG2 ? 2 : 1);       // CG2: both CTAs' producers arrive (with their byte counts) on the leader's
      mbar_init(&empty_bar[s], (CL2 && !CG2) ? 2 : 1);      // CL2: released by both CTAs' MMA warps (multicast commit)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CG2 ? 16 : 8);    // CG2: the leader's MMA waits for both CTAs' epilogue warps
    }
    for (int s = 0; s < SLOTS; ++s) mbar_init(&rfull_bar[s], 1);
    if constexpr (XF)
      for (int s = 0; s < STAGES; ++s) mbar_init(&xf_bar[s], 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG2) {
      tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL2) cluster_sync_all();        // the peer's barriers exist before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                   // everything above overlapped the previous kernel's tail (PDL)

  if (warp == 0) {
    if (lane == 0) gemm_producer<BN, CL2, CG2, XF>(tmA0, tmA1, tmA2, tmB, a, smem_a, smem_b, full_bar, empty_bar);
  } else if (warp == 1) {
    if (lane == 0 && (!CG2 || cluster_ctarank() == 0))      // CG2: only the pair's leader issues MMAs
      gemm_mma<BN, CL2, CG2, XF>(a, smem_a, smem_b, XF ? xf_bar : full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base);
  } else if (XF && warp >= 10) {
    gemm_transform<BN>(a, smem_a, xf_tab, full_bar, xf_bar);
  } else {
    // ===================== epilogue (warps 2..9, 256 threads) =====================
    const int et = threadIdx.x - 64;        // 0..255
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which 32 columns of the 64-column sub-box
    const int row = q * 32 + lane;          // row inside the 128-row tile
    const bool is_e = (et == 0);            // elected thread: TMA stores / residual loads / bulk-group bookkeeping
    const uint32_t slot0 = smem_u32(smem_slot);
    const uint32_t bias0 = smem_u32(smem_bias);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t rsw = static_cast<uint32_t>(row & 7);

    // accumulator buffer `ac` fully read: tell the MMA issuer (CG2: the pair's leader, possibly in the other CTA)
    auto arrive_tempty = [&](int ac) {
      if constexpr (CG2) {
        if (cluster_ctarank() == 0)
          mbar_arrive(&tempty_bar[ac]);
        else
          mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[ac]), 0));
      } else {
        mbar_arrive(&tempty_bar[ac]);
      }
    };
    // sub-box kk of this CTA's tile sequence -> global coordinates; issues the residual TMA load
    auto issue_res = [&](uint32_t kk) {
      int mt, nt;
      if (!gemm_tile_at<CL2>(a, static_cast<int>(kk / NSUB), mt, nt)) return;
      const int j = kk % NSUB;
      const uint32_t s = kk % SLOTS;
      mbar_arrive_expect_tx(&rfull_bar[s], Cfg::SLOT_BYTES);
      tma_load_2d(smem_slot + s * Cfg::SLOT_BYTES, &tmR, &rfull_bar[s], nt * BN + j * 64, mt * 128);
    };
    if (is_e && has_res) {
      issue_res(0);
      issue_res(1);
    }

    bool num_tiles_done = false;
    uint32_t k = 0;                         // sub-box counter of this CTA
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (NSUB % 2 == 0) {
      if (a.geglu) {
        // ---- GEGLU epilogue: input sub-boxes (2jj, 2jj+1) = (hidden, gate) columns of the same 64 output features
        int mt, nt;
        for (int it = 0; gemm_tile_at<CL2>(a, it, mt, nt); ++it) {
          if (et < BN) {
            const int n = nt * BN + et;
            smem_bias[acc * BN + et] = (a.bias != nullptr && n < a.N) ? __ldg(a.bias + n) : 0.f;
          }
          named_barrier_sync(1, 256);
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
          for (int jj = 0; jj < NSUB / 2; ++jj, ++k) {
            const uint32_t slot = slot0 + (k % SLOTS) * Cfg::SLOT_BYTES;
            uint32_t rh[32], rg[32];
            tmem_ld32(t_row + (2 * jj) * 64 + half * 32, rh);
            tmem_ld32(t_row + (2 * jj + 1) * 64 + half * 32, rg);
            tmem_ld_wait();
            if (jj == NSUB / 2 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) arrive_tempty(acc);
            }
            const uint32_t bh = bias0 + static_cast<uint32_t>(acc * BN + (2 * jj) * 64 + half * 32) * 4u;
            const uint32_t bg = bh + 64u * 4u;
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 b1 = lds128(bh + i * 16), b2 = lds128(bg + i * 16);
              const float hb[4] = {__uint_as_float(b1.x), __uint_as_float(b1.y), __uint_as_float(b1.z), __uint_as_float(b1.w)};
              const float gb[4] = {__uint_as_float(b2.x), __uint_as_float(b2.y), __uint_as_float(b2.z), __uint_as_float(b2.w)};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                // the unfused path rounds the projection to bf16 before GEGLU: keep that rounding so that both paths agree
                const float hv = act2f(f2act(__uint_as_float(rh[4 * i + c]) + hb[c]));
                const float gv = act2f(f2act(__uint_as_float(rg[4 * i + c]) + gb[c]));
                v[4 * i + c] = hv * gelu_erf(gv);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = pack_act(v[8 * i + 0], v[8 * i + 1]);
              u.y = pack_act(v[8 * i + 2], v[8 * i + 3]);
              u.z = pack_act(v[8 * i + 4], v[8 * i + 5]);
              u.w = pack_act(v[8 * i + 6], v[8 * i + 7]);
              sts128(slot + row_off + (((half * 4 + i) ^ rsw) << 4), u);
            }
            fence_proxy_async_smem();
            if (is_e) bulk_wait_group_read<1>();      // stores <= k-3 have released their slots (4-slot ring)
            named_barrier_sync(1, 256);
            if (is_e) {
              tma_store_2d(&tmO, smem_slot + (k % SLOTS) * Cfg::SLOT_BYTES, nt * (BN / 2) + jj * 64, mt * 128);
              bulk_commit_group();
            }
          }
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
        if (is_e) bulk_wait_group<0>();
        num_tiles_done = true;
      }
    }
    int mt, nt;
    for (int it = 0; !num_tiles_done && gemm_tile_at<CL2>(a, it, mt, nt); ++it) {
      // bias of this tile's columns -> smem (double buffered by accumulator index)
      if (et < BN) {
        const int n = nt * BN + et;
        smem_bias[acc * BN + et] = (a.bias != nullptr && n < a.N) ? __ldg(a.bias + n) : 0.f;
      }
      named_barrier_sync(1, 256);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      uint32_t r[32];
      tmem_ld32(t_row + half * 32, r);
#pragma unroll 1
      for (int j = 0; j < NSUB; ++j, ++k) {
        const uint32_t slot = slot0 + (k % SLOTS) * Cfg::SLOT_BYTES;
        if (has_res) mbar_wait(&rfull_bar[k % SLOTS], (k / SLOTS) & 1);
        tmem_ld_wait();
        float v[32];
        const uint32_t bsrc = bias0 + static_cast<uint32_t>(acc * BN + j * 64 + half * 32) * 4u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 bb = lds128(bsrc + i * 16);
          v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + __uint_as_float(bb.x);
          v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + __uint_as_float(bb.y);
          v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + __uint_as_float(bb.z);
          v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + __uint_as_float(bb.w);
        }
        if (j == NSUB - 1) {                // accumulator fully read: hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_tempty(acc);
        } else {                            // next sub-box's accumulator columns stream in behind this one's maths
          tmem_ld32(t_row + (j + 1) * 64 + half * 32, r);
        }
        if (a.act == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __fdividef(v[i], 1.f + __expf(-1.702f * v[i]));
        }
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = lds128(slot + row_off + (((half * 4 + i) ^ rsw) << 4));
            float2 f;
            f = unpack_act(u.x); v[8 * i + 0] += f.x; v[8 * i + 1] += f.y;
            f = unpack_act(u.y); v[8 * i + 2] += f.x; v[8 * i + 3] += f.y;
            f = unpack_act(u.z); v[8 * i + 4] += f.x; v[8 * i + 5] += f.y;
            f = unpack_act(u.w); v[8 * i + 6] += f.x; v[8 * i + 7] += f.y;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_act(v[8 * i + 0] * a.out_scale, v[8 * i + 1] * a.out_scale);
          u.y = pack_act(v[8 * i + 2] * a.out_scale, v[8 * i + 3] * a.out_scale);
          u.z = pack_act(v[8 * i + 4] * a.out_scale, v[8 * i + 5] * a.out_scale);
          u.w = pack_act(v[8 * i + 6] * a.out_scale, v[8 * i + 7] * a.out_scale);
          sts128(slot + row_off + (((half * 4 + i) ^ rsw) << 4), u);
        }
        fence_proxy_async_smem();           // generic-proxy writes -> visible to the TMA store
        if (is_e) {
          // stores up to sub-box k-2 have finished reading their slots (k-1 may still be in flight), and every
          // thread finished its statistics reads of k-2 before barrier k-1: slot (k+2)%4 == (k-2)%4 is free
          bulk_wait_group_read<1>();
          if (has_res) issue_res(k + 2);
        }
        named_barrier_sync(1, 256);
        if (is_e) {
          if (a.out4d) {                     // one phase of the fused upsample: strided pixels of the high-res tensor
            int n0, y0 = 0, x0 = 0;
            if (a.tiles_per_img > 0) {
              n0 = mt / a.tiles_per_img;
              const int r = mt % a.tiles_per_img;
              y0 = (r / a.x_chunks) * a.tileH;
              x0 = (r % a.x_chunks) * 128;
            } else {
              n0 = mt * a.tileN;
            }
            tma_store_4d(&tmO, smem_slot + (k % SLOTS) * Cfg::SLOT_BYTES, nt * BN + j * 64, x0, y0, n0);
          } else {
            tma_store_2d(&tmO, smem_slot + (k % SLOTS) * Cfg::SLOT_BYTES, nt * BN + j * 64, mt * 128);
          }
          bulk_commit_group();
        }
        if (a.gn_stats != nullptr) {
          // column sums of the stored bf16 values, fixed order, no atomics.  lane -> (row residue rg = lane & 7,
          // column pair cp = lane >> 3) of the warp's 16-byte unit; rows rg + 8*i: the 8 residues hit 8 different
          // swizzle positions (conflict-free) and the per-row offset i*1024 is an immediate.
          const uint32_t rg = lane & 7, cp = lane >> 3;
          const uint32_t unit = static_cast<uint32_t>(et >> 5);
          const uint32_t base = slot + rg * 128u + ((unit ^ rg) << 4) + cp * 4u;
          float s0[2] = {0.f, 0.f}, q0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f}, q1[2] = {0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t w = lds32(base + i * 1024u);
            const float2 xx = unpack_act(w);
            const float x0 = xx.x, x1 = xx.y;
            const int hf = i >> 3;
            s0[hf] += x0;
            q0[hf] = fmaf(x0, x0, q0[hf]);
            s1[hf] += x1;
            q1[hf] = fmaf(x1, x1, q1[hf]);
          }
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              s0[hf] += __shfl_xor_sync(0xffffffffu, s0[hf], o);
              q0[hf] += __shfl_xor_sync(0xffffffffu, q0[hf], o);
              s1[hf] += __shfl_xor_sync(0xffffffffu, s1[hf], o);
              q1[hf] += __shfl_xor_sync(0xffffffffu, q1[hf], o);
            }
          }
          if (rg < 2) {                               // lane rg = 0 writes the first 64-row half, rg = 1 the second
            const int hrow_in = mt * 2 + static_cast<int>(rg);
            // phase launches of the fused upsample: the statistics rows of one image are the 4 phases' rows back to back
            const int hrow = a.stats_in_rows > 0
                                 ? (hrow_in / a.stats_in_rows) * a.stats_img_rows + a.stats_off + hrow_in % a.stats_in_rows
                                 : hrow_in;
            const int n = nt * BN + j * 64 + static_cast<int>(unit * 8 + cp * 2);
            if (hrow_in * 64 < a.M && n + 1 < a.N) {
              const float4 o4 = rg == 0 ? make_float4(s0[0], q0[0], s1[0], q1[0]) : make_float4(s0[1], q0[1], s1[1], q1[1]);
              *reinterpret_cast<float4*>(a.gn_stats + static_cast<size_t>(hrow) * a.ld_stats + n) = o4;
            } else if (hrow_in * 64 < a.M && n < a.N) {
              a.gn_stats[static_cast<size_t>(hrow) * a.ld_stats + n] = rg == 0 ? make_float2(s0[0], q0[0]) : make_float2(s0[1], q0[1]);
            }
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (is_e) bulk_wait_group<0>();         // all stores complete before shared memory goes away
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL2) cluster_sync_all();        // nobody leaves while the peer may still multicast / arrive into this CTA
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG2)
      tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    else
      tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// BN = 16 kernel (N <= 16 per tile): fp32 or bf16 output, plain per-row global stores.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(320, 1)
gemm_small_n_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB, const GemmArgs a) {
  constexpr int BN = 16;
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  pdl_launch_dependents();
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
  pdl_wait();                   // everything above overlapped the previous kernel's tail (PDL)

  if (warp == 0) {
    if (lane == 0) gemm_producer<BN>(tmA0, tmA1, tmA2, tmB, a, smem_a, smem_b, full_bar, empty_bar);
  } else if (warp == 1) {
    if (lane == 0) gemm_mma<BN>(a, smem_a, smem_b, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base);
  } else if (warp < 6) {
    // epilogue: warps 2..5, one per TMEM lane quarter
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool has_res = a.residual != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tl = a.reverse ? num_tiles - 1 - tile : tile;
    const int mt = tl / a.n_tiles, nt = tl % a.n_tiles;
      const int m = mt * 128 + row;
      const bool m_ok = m < a.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      const int n_base = nt * BN;
      if (m_ok) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n_base + j;
          if (n < a.N) {
            float v = __uint_as_float(r[j]);
            if (a.bias != nullptr) v += __ldg(a.bias + n);
            if (a.act == 1) v = __fdividef(v, 1.f + __expf(-1.702f * v));
            if (has_res) v += act2f(a.residual[static_cast<size_t>(m) * a.ld_res + n]);
            v *= a.out_scale;
            if (a.out_fp32)
              reinterpret_cast<float*>(a.out)[static_cast<size_t>(m) * a.ld_out + n] = v;
            else
              reinterpret_cast<act_t*>(a.out)[static_cast<size_t>(m) * a.ld_out + n] = f2act(v);
          }
        }
      }
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
