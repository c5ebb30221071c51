// Exact baseline-JPEG byte count of uint8 RGB images on the GPU: the compressibility scorer
// (edm/scorers.py:176-244 calls PIL -> libjpeg-turbo: quality 80, 4:2:0, standard Huffman tables).
// One CTA per image, one thread per 8x8 block: libjpeg's integer pipeline (jccolor.c rgb_ycc_convert,
// jcsample.c h2v2_downsample, jfdctint.c jpeg_fdct_islow, jcdctmgr.c quantisation, jchuff.c
// encode_one_block) followed by a real bit stream in shared memory so that the 0xFF byte stuffing and
// the final 1-bit padding are counted exactly.  Integer arithmetic only: bit-exact against libjpeg.
#pragma once
#include "common.cuh"

namespace b200 {

struct JpegTables {          // built on the host from a header libjpeg itself wrote (tables + header length)
  int32_t q[2][64];          // quantisation tables, natural order (0 = luminance, 1 = chrominance)
  int32_t dc_len[2][16];
  int32_t dc_code[2][16];
  int32_t ac_len[2][256];
  int32_t ac_code[2][256];
  int32_t header_bytes;      // bytes up to and including the SOS header
};

__constant__ int c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

DEVINL int jdesc(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jfdctint.c: one 1-D pass over v[0..7] (stride s); first pass scales up by 2^PASS1_BITS
DEVINL void fdct_pass(int* v, int s, bool first) {
  constexpr int CB = 13, P1 = 2;
  constexpr int F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299,
                F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
  int t0 = v[0] + v[7 * s], t7 = v[0] - v[7 * s];
  int t1 = v[s] + v[6 * s], t6 = v[s] - v[6 * s];
  int t2 = v[2 * s] + v[5 * s], t5 = v[2 * s] - v[5 * s];
  int t3 = v[3 * s] + v[4 * s], t4 = v[3 * s] - v[4 * s];
  const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  const int sh = first ? CB - P1 : CB + P1;
  v[0] = first ? (t10 + t11) << P1 : jdesc(t10 + t11, P1);
  v[4 * s] = first ? (t10 - t11) << P1 : jdesc(t10 - t11, P1);
  int z1 = (t12 + t13) * F0_541;
  v[2 * s] = jdesc(z1 + t13 * F0_765, sh);
  v[6 * s] = jdesc(z1 - t12 * F1_847, sh);
  z1 = t4 + t7;
  int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
  const int z5 = (z3 + z4) * F1_175;
  t4 *= F0_298; t5 *= F2_053; t6 *= F3_072; t7 *= F1_501;
  z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
  z3 += z5; z4 += z5;
  v[7 * s] = jdesc(t4 + z1 + z3, sh);
  v[5 * s] = jdesc(t5 + z2 + z4, sh);
  v[3 * s] = jdesc(t6 + z2 + z3, sh);
  v[s] = jdesc(t7 + z1 + z4, sh);
}

DEVINL int nbits_of(int v) { return v == 0 ? 0 : 32 - __clz(v < 0 ? -v : v); }

struct BitSink {            // either counts bits or ORs them into a zeroed big-endian word stream
  uint32_t* words;
  int pos;
  DEVINL void put(int code, int n) {
    if (n == 0) return;
    if (words != nullptr) {
      const uint32_t c = static_cast<uint32_t>(code) & ((1u << n) - 1u);
      const int w = pos >> 5, sh = 32 - (pos & 31) - n;
      if (sh >= 0) {
        atomicOr(&words[w], c << sh);
      } else {
        atomicOr(&words[w], c >> (-sh));
        atomicOr(&words[w + 1], c << (32 + sh));
      }
    }
    pos += n;
  }
};

// jchuff.c encode_one_block on zig-zag ordered quantised coefficients zz[0..63] (stride 1)
DEVINL void encode_block(const short* zz, int pred, const JpegTables& T, int tab, BitSink& sink) {
  const int diff = zz[0] - pred;
  int n = nbits_of(diff);
  sink.put(T.dc_code[tab][n], T.dc_len[tab][n]);
  sink.put(diff >= 0 ? diff : diff - 1, n);
  int run = 0;
  for (int k = 1; k < 64; ++k) {
    const int v = zz[k];
    if (v == 0) {
      ++run;
      continue;
    }
    while (run > 15) {
      sink.put(T.ac_code[tab][0xF0], T.ac_len[tab][0xF0]);
      run -= 16;
    }
    n = nbits_of(v);
    sink.put(T.ac_code[tab][(run << 4) | n], T.ac_len[tab][(run << 4) | n]);
    sink.put(v >= 0 ? v : v - 1, n);
    run = 0;
  }
  if (run) sink.put(T.ac_code[tab][0], T.ac_len[tab][0]);
}

// img uint8 [M,3,H,W] (H, W multiples of 16, <= 64) -> sizes int32 [M], scores fp32 [M].
// dynamic smem: see jpeg_smem_bytes().
__global__ void jpeg_size_kernel(const uint8_t* __restrict__ img, const JpegTables* __restrict__ tables, int H, int W,
                                 float min_size, float max_size, int32_t* __restrict__ sizes,
                                 float* __restrict__ scores) {
  extern __shared__ uint8_t smem_raw[];
  const int HW = H * W, Hc = H / 2, Wc = W / 2;
  const int nby = H / 8, nbx = W / 8, nY = nby * nbx, nC = (Hc / 8) * (Wc / 8), nb = nY + 2 * nC;
  const int mcus_x = W / 16, n_mcu = (H / 16) * mcus_x;
  JpegTables* T = reinterpret_cast<JpegTables*>(smem_raw);
  short* plane = reinterpret_cast<short*>(smem_raw + ((sizeof(JpegTables) + 15) & ~15));      // Y [H*W], Cb, Cr [Hc*Wc]
  int* ws = reinterpret_cast<int*>(plane + HW + 2 * Hc * Wc);                                 // [nb][65] workspace
  short* zz = reinterpret_cast<short*>(ws + nb * 65);                                        // [nb][64] zig-zag quantised
  int* blk_bits = reinterpret_cast<int*>(zz + nb * 64);                                      // [nb + 1] (coding order)
  uint32_t* stream = reinterpret_cast<uint32_t*>(blk_bits + nb + 2);                         // [nb * 64] words
  __shared__ int s_ff;

  const int tid = threadIdx.x;
  const uint8_t* im = img + static_cast<size_t>(blockIdx.x) * 3 * HW;
  for (int i = tid; i < static_cast<int>(sizeof(JpegTables) / 4); i += blockDim.x)
    reinterpret_cast<int*>(T)[i] = reinterpret_cast<const int*>(tables)[i];
  for (int i = tid; i < nb * 64; i += blockDim.x) stream[i] = 0u;
  if (tid == 0) s_ff = 0;
  // ---- colour conversion (jccolor.c) + h2v2 chroma down-sampling (jcsample.c)
  for (int p = tid; p < HW; p += blockDim.x) {
    const int r = im[p], g = im[HW + p], b = im[2 * HW + p];
    plane[p] = static_cast<short>(((19595 * r + 38470 * g + 7471 * b + 32768) >> 16) - 128);
  }
  for (int p = tid; p < Hc * Wc; p += blockDim.x) {
    const int cy = p / Wc, cx = p - cy * Wc;
    int sb = 0, sr = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int q = (2 * cy + (t >> 1)) * W + 2 * cx + (t & 1);
      const int r = im[q], g = im[HW + q], b = im[2 * HW + q];
      sb += (-11059 * r - 21709 * g + 32768 * b + 8388608 + 32767) >> 16;
      sr += (32768 * r - 27439 * g - 5329 * b + 8388608 + 32767) >> 16;
    }
    const int bias = 1 + (cx & 1);
    plane[HW + p] = static_cast<short>(((sb + bias) >> 2) - 128);
    plane[HW + Hc * Wc + p] = static_cast<short>(((sr + bias) >> 2) - 128);
  }
  __syncthreads();
  // ---- per-block FDCT + quantisation; block id: [0,nY) = Y raster, then Cb raster, then Cr raster
  for (int blk = tid; blk < nb; blk += blockDim.x) {
    const short* src;
    int pw, by, bx, tab;
    if (blk < nY) {
      src = plane; pw = W; by = blk / nbx; bx = blk - by * nbx; tab = 0;
    } else {
      const int c = blk - nY, comp = c / nC, cc = c - comp * nC;
      src = plane + HW + comp * Hc * Wc; pw = Wc; by = cc / (Wc / 8); bx = cc - by * (Wc / 8); tab = 1;
    }
    int* w = ws + blk * 65;
    for (int y = 0; y < 8; ++y)
      for (int x = 0; x < 8; ++x) w[y * 8 + x] = src[(by * 8 + y) * pw + bx * 8 + x];
    for (int y = 0; y < 8; ++y) fdct_pass(w + y * 8, 1, true);
    for (int x = 0; x < 8; ++x) fdct_pass(w + x, 8, false);
    short* z = zz + blk * 64;
    for (int k = 0; k < 64; ++k) {
      const int nat = c_zigzag[k];
      const int qv = T->q[tab][nat] << 3;
      const int cf = w[nat];
      const int a = ((cf < 0 ? -cf : cf) + (qv >> 1)) / qv;
      z[k] = static_cast<short>(cf < 0 ? -a : a);
    }
  }
  __syncthreads();
  // ---- coding order: MCU m = 4 Y blocks (2x2), Cb, Cr; DC predictor = previous block of the same component
  auto seq_to_blk = [&](int seq, int& comp) {
    const int m = seq / 6, k = seq - m * 6;
    const int my = m / mcus_x, mx = m - my * mcus_x;
    if (k < 4) {
      comp = 0;
      return (my * 2 + (k >> 1)) * nbx + mx * 2 + (k & 1);
    }
    comp = k - 3;
    return nY + (comp - 1) * nC + my * (Wc / 8) + mx;
  };
  auto dc_pred = [&](int seq) {
    const int m = seq / 6, k = seq - m * 6;
    int comp, prev_seq;
    if (k >= 1 && k < 4) prev_seq = seq - 1;
    else if (m == 0) return 0;
    else prev_seq = (k == 0) ? (m - 1) * 6 + 3 : seq - 6;
    return static_cast<int>(zz[seq_to_blk(prev_seq, comp) * 64]);
  };
  for (int seq = tid; seq < nb; seq += blockDim.x) {
    int comp;
    const int blk = seq_to_blk(seq, comp);
    BitSink cnt{nullptr, 0};
    encode_block(zz + blk * 64, dc_pred(seq), *T, comp ? 1 : 0, cnt);
    blk_bits[seq] = cnt.pos;
  }
  __syncthreads();
  if (tid == 0) {                      // exclusive prefix sum over <= 96 blocks
    int acc = 0;
    for (int s = 0; s < nb; ++s) {
      const int l = blk_bits[s];
      blk_bits[s] = acc;
      acc += l;
    }
    blk_bits[nb] = acc;
  }
  __syncthreads();
  for (int seq = tid; seq < nb; seq += blockDim.x) {
    int comp;
    const int blk = seq_to_blk(seq, comp);
    BitSink out{stream, blk_bits[seq]};
    encode_block(zz + blk * 64, dc_pred(seq), *T, comp ? 1 : 0, out);
  }
  __syncthreads();
  const int total_bits = blk_bits[nb];
  const int nbytes = (total_bits + 7) >> 3;
  if (tid == 0 && (total_bits & 7)) {                    // pad the last byte with 1-bits (jchuff.c flush_bits)
    const int pad = 8 - (total_bits & 7);
    BitSink out{stream, total_bits};
    out.put((1 << pad) - 1, pad);
  }
  __syncthreads();
  int ff = 0;
  for (int i = tid; i < nbytes; i += blockDim.x) ff += ((stream[i >> 2] >> (24 - 8 * (i & 3))) & 0xFFu) == 0xFFu;
  if (ff) atomicAdd(&s_ff, ff);
  __syncthreads();
  if (tid == 0) {
    const int size = T->header_bytes + nbytes + s_ff + 2;             // + EOI
    sizes[blockIdx.x] = size;
    const double nrm = (static_cast<double>(size) - min_size) / (static_cast<double>(max_size) - min_size);
    scores[blockIdx.x] = static_cast<float>(1.0 - fmin(1.0, fmax(0.0, nrm)));
  }
}

inline size_t jpeg_smem_bytes(int H, int W) {
  const int HW = H * W, nb = (H / 8) * (W / 8) * 3 / 2;
  return ((sizeof(JpegTables) + 15) & ~size_t(15)) + sizeof(short) * (HW + HW / 2) + sizeof(int) * nb * 65 +
         sizeof(short) * nb * 64 + sizeof(int) * (nb + 2) + sizeof(uint32_t) * nb * 64 + 64;
}

}  // namespace b200
