/* b200_noise_search.h -- C ABI of libb200ns.so: the sm_100a kernels behind the
 * candidate-batched noise-search step of rvignav/diffusion-tts.
 *
 * The reference has no FFI of its own (pure Python/PyTorch): the seam is the
 * Python call sites listed per entry point below (paths relative to the reference checkout).
 * Every pointer is a DEVICE pointer unless stated; every function enqueues on `stream`
 * (a cudaStream_t passed as void*) and returns 0 on success or a non-zero code, with the
 * message available from b200ns_last_error().  There is no CPU fallback anywhere.
 *
 * Layout conventions
 *   sampler state      fp64, NCHW flattened to [R, E], R = N*b candidate rows (row = n*b + j),
 *                      E = C*H*W                               (edm/main.py:803-806 layout)
 *   U-Net activations  bf16 NHWC [B, H, W, C], C a multiple of 64
 *   U-Net weights      bf16 [Cout_pad, Ktot], K-major; K order = (segment, tap kh*3+kw, channel)
 */
#ifndef B200_NOISE_SEARCH_H
#define B200_NOISE_SEARCH_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* b200ns_last_error(void);
/* 1 if the library stores activations / weights as IEEE half (the default build), 0 for the bfloat16 build
 * (-DB200NS_ACT_BF16).  "bf16" in the comments below means this 16-bit storage type. */
int b200ns_act_is_fp16(void);
/* 1 if device `dev` is compute capability 10.x (B200), else 0. */
int b200ns_device_ok(int dev);

/* ------------------------------------------------------------------ sampler / scorer
 * x_hat = x_cur + s*eps ; net_in = c_in * fp32(x_hat)                 edm/main.py:85,
 * edm/training/networks.py:655,665.  x_cur is [b,E] (broadcast over candidates: row % b),
 * eps/x_hat are [R,E] fp64, net_in is fp32 [R,E] (NCHW).  s = sqrt(t_hat^2-t_cur^2)*S_noise. */
int b200ns_heun_pre(const double* x_cur, const double* eps, double* x_hat, float* net_in,
                    int64_t R, int64_t b, int64_t E, double s, float c_in, void* stream);
/* The same with an fp32 noise tensor (the MCTS depth noises edm/main.py:445, or an fp32 precomputed_noise): torch evaluates
 * `sqrt(...) * S_noise * eps_i` with a 0-dim fp64 scale and an fp32 tensor as an fp32 product (scale rounded to fp32), and only
 * the sum with x_cur in fp64:  x_hat = x_cur + double(fp32(s) * eps). */
int b200ns_heun_pre_f32noise(const double* x_cur, const float* eps, double* x_hat, float* net_in,
                             int64_t R, int64_t b, int64_t E, double s, float c_in, void* stream);

/* Euler half step: D1 = c_skip*fp32(x_hat) + c_out*F1 ; d = (x_hat - D1)/t_hat ;
 * x_eul = x_hat + dt*d ; net_in2 = c_in_next*fp32(x_eul).            edm/main.py:87-91,
 * networks.py:667.  F1 is the U-Net output, fp32 NHWC [R, HW, C].  x_eul may be NULL. */
int b200ns_heun_mid(const double* x_hat, const float* F1, float* net_in2, double* x_eul,
                    int64_t R, int32_t C, int32_t HW, float c_skip, float c_out, double t_hat,
                    double dt, float c_in_next, void* stream);

/* Second-order correction + Tweedie x0 + uint8 quantise + per-channel integer pixel sums.
 * F2 == NULL selects the last step (no correction; x_next = x_eul, denoised = D1).
 * Outputs (each may be NULL): x_next fp64 [R,E]; x0_u8 uint8 [R,C,HW];
 * chan_sums uint32 [R,4] (zeroed by this call, then accumulated with integer atomics).
 * edm/main.py:88-96, 825-827; edm/scorers.py:37-46. */
int b200ns_heun_post(const double* x_hat, const float* F1, const float* F2, double* x_next,
                     uint8_t* x0_u8, uint32_t* chan_sums, int64_t R, int32_t C, int32_t HW,
                     float c_skip1, float c_out1, double t_hat, double dt, float c_skip2,
                     float c_out2, double t_next, void* stream);

/* (x*127.5+128).clip(0,255) truncated to uint8.                      edm/main.py:827,869 */
int b200ns_quantize_u8(const double* x, uint8_t* out, int64_t n, void* stream);

/* Per-channel integer pixel sums of uint8 images [M,C,HW] -> uint32 [M,4] (C <= 4). */
int b200ns_channel_sums_u8(const uint8_t* img, uint32_t* chan_sums, int64_t M, int32_t C,
                           int32_t HW, void* stream);

/* BrightnessScorer from integer sums: C==3 -> clamp(sum_c w_c*S_c/(255*HW),0,1) with
 * w=(0.2126,0.7152,0.0722); otherwise the plain mean over C*HW.      edm/scorers.py:37-52,
 * sd/scorers.py:66-67. */
int b200ns_brightness_from_sums(const uint32_t* chan_sums, float* scores, int64_t M, int32_t C,
                                int32_t HW, void* stream);

/* First-maximal argmax over candidates: scores [N,b] -> idx [b] (int64, local index n) and the
 * key (orderable_i32(score) << 32) | (0xFFFFFFFF - (idx_base + n)) [b] (may be NULL), a SIGNED
 * int64 whose max picks the best score and, among equal scores, the lowest global index -- the
 * same rule as torch.argmax -- so shards combine with ncclAllReduce(max, int64).
 * edm/main.py:842. */
int b200ns_argmax_first(const float* scores, int64_t N, int64_t b, int64_t idx_base, int64_t* idx,
                        int64_t* packed_key, void* stream);

/* dst[j,:] = src[idx[j], j, :] for src [N,b,E] fp64.                  edm/main.py:848-851 */
int b200ns_gather_rows(const double* src, const int64_t* idx, double* dst, int64_t N, int64_t b,
                       int64_t E, void* stream);

/* ||dirs[r,:]||_2 in fp64 (deterministic block reduction).           edm/main.py:764,770 */
int b200ns_direction_norms(const double* dirs, double* norms, int64_t R, int64_t E, void* stream);

/* cand[r] = fresh_mask[r] ? fresh[r] : pivot[r % b] + (double)scale[r] * (dirs[r]/norms[r]).
 * edm/main.py:749-800.  fresh may be NULL when no row is masked. */
int b200ns_make_candidates(const double* pivot, const double* dirs, const double* norms,
                           const float* scale, const uint8_t* fresh_mask, const double* fresh,
                           double* cand, int64_t R, int64_t b, int64_t E, void* stream);

/* Exact baseline-JPEG byte count (PIL/libjpeg: 4:2:0, standard Huffman tables) and the compressibility
 * score 1 - clip((size-min)/(max-min),0,1) of uint8 RGB images [M,3,H,W], H and W multiples of 16, <= 64.
 * `tables` is a device copy of the table block (b200ns_jpeg_tables_bytes() bytes, int32 fields: q[2][64],
 * dc_len[2][16], dc_code[2][16], ac_len[2][256], ac_code[2][256], header_bytes) built from a header that
 * libjpeg itself wrote for this size/quality.                         edm/scorers.py:207-244 */
int b200ns_jpeg_size(const uint8_t* img, const void* tables, int64_t M, int32_t H, int32_t W, float min_size,
                     float max_size, int32_t* sizes, float* scores, void* stream);
int b200ns_jpeg_tables_bytes(void);

/* ------------------------------------------------------------------ U-Net engine (plans)
 * A plan is an ordered list of kernel launches with all shapes, pointers and TMA descriptors
 * resolved at build time; b200ns_plan_run enqueues them on one stream.  It is the
 * B200 replacement for DhariwalUNet.forward / SongUNet.forward
 * (edm/training/networks.py:435-461, 320-363) and their leaves (:39-43, 68-90, 104-106,
 * 115-118, 166-187). */
typedef struct b200ns_plan b200ns_plan;
b200ns_plan* b200ns_plan_create(void);
void b200ns_plan_destroy(b200ns_plan* p);
/* Programmatic dependent launch for this plan's kernels (call before b200ns_plan_instantiate_graph): -1 = the process-wide
 * B200NS_PDL setting, 0 off, 1 every kernel, 2 GroupNorm kernels only.  Batch-1..4 plans are latency-bound (one 32x32 DDPM++
 * NFE at batch 1: 1.83 -> 1.69 ms with mode 1); large batches are power-bound and lose ~1 % (profiles/r01_pdl_ab.txt). */
int b200ns_plan_set_pdl(b200ns_plan* p, int mode);
int b200ns_plan_size(const b200ns_plan* p);
/* Valid output columns of GEMM launch `op` (-1 if `op` is not a GEMM).  One b200ns_plan_add_gemm call may append two
 * launches over column slices (see there), so callers that keep per-launch metadata ask how the columns were split. */
int b200ns_plan_gemm_cols(const b200ns_plan* p, int op);
int b200ns_plan_run(b200ns_plan* p, void* stream);
/* Capture the plan once as a CUDA graph; later b200ns_plan_run calls launch the graph. */
int b200ns_plan_instantiate_graph(b200ns_plan* p);
/* Ops added after this call belong to `lane` (0..3).  Lane 0 is the main stream; in the captured graph the ops
 * of a lane k > 0 form a parallel branch (forked at the lane's first op, joined before the next lane-0 op), so
 * HBM-bound kernels of one half batch overlap the tensor-core kernels of the other.  Eager runs ignore lanes. */
int b200ns_plan_set_lane(b200ns_plan* p, int lane);
/* Run ops [first, last) only (per-layer parity tests and profiling). */
int b200ns_plan_run_range(b200ns_plan* p, int first, int last, void* stream);

typedef struct {
  int32_t src;      /* 0..2: which activation tensor */
  int32_t taps;     /* 1 (1x1 / plain GEMM) or 9 (3x3, pad 1) */
  int32_t cstart;   /* first channel inside the source (multiple of 64) */
  int32_t cblocks;  /* number of 64-channel blocks */
} b200ns_kseg;

/* Implicit-GEMM convolution / GEMM on tcgen05 tensor cores:
 *   out[m, n] = out_scale * ( sum_k A[m,k]*W[n,k] + bias[n] + residual[m,n] )
 * Replaces Conv2d.forward (networks.py:68-90) for kernel 3 and 1, the channel concat of
 * the decoder (networks.py:458) via several activation sources, and the fused
 * `conv1(x) + skip(orig)` of UNetBlock.forward (networks.py:177-179) via several K segments. */
typedef struct {
  const void* a_ptr[3];      /* bf16 NHWC [batch,H,W,a_channels[i]]; unused = NULL */
  int32_t a_channels[3];
  int32_t n_seg;
  b200ns_kseg seg[8];
  int32_t batch, H, W;
  const void* w_ptr;         /* bf16 [Npad, Ktot] */
  int32_t N, Npad, Ktot;
  const float* bias;         /* [N] or NULL */
  const void* residual;      /* bf16 [M, ld_res] or NULL */
  int32_t ld_res;
  float out_scale;
  void* out;                 /* bf16 or fp32 [M, ld_out] */
  int32_t ld_out;
  int32_t out_fp32;
  /* optional (bf16 output, M % 64 == 0): fp32 [M/64, N, 2] = per-channel (sum, sum of squares) of the
   * STORED values over each 64-row half tile, reduced in the epilogue in a fixed order; feeds
   * b200ns_plan_add_gn_finalize so the consumer's GroupNorm needs no statistics pass over HBM. */
  float* gn_stats;
  /* 1: process the output tiles last-to-first.  Alternating the walk direction between producer and consumer
   * makes the consumer start on the rows the producer wrote last, which are still in the 126 MB L2. */
  int32_t reverse;
  /* per source: 1, or 2 = the source is [batch, 2H, 2W, C] and is sampled with stride 2, i.e. a 3x3 stride-2 pad-1
   * convolution (Downsample2D, sd/diffusers/src/diffusers/models/downsampling.py) through TMA elementStrides. */
  int32_t a_stride[3];
  /* 1: GEGLU fused into the epilogue (diffusers GEGLU.forward, activations.py:117-123).  The weight rows / bias entries
   * come in groups of 128 = [64 hidden rows | 64 gate rows] of the same 64 output features; `out` is bf16 [M, N/2]
   * (ld_out >= N/2) = hidden * gelu(gate), exact erf GELU.  No residual / gn_stats; Npad % 128 == 0. */
  int32_t geglu;
  /* 1: out [batch, 2H, 2W, ld_out] = conv3x3(nearest_upsample_2x(A)) (Conv2d(up=True), networks.py:72-80 with
   * resample_filter [1,1]; diffusers Upsample2D) WITHOUT materialising the upsampled tensor: H, W are the LOW-res dims of
   * the A tensors, every segment says taps = 9, w_ptr = bf16 [4][Npad][Ktot] with Ktot = 4 * 64 * sum(cblocks): per output
   * phase (py, px) the 3x3 weights that read the same source pixel pre-summed into a 2x2 kernel (tap order a*2+c; source
   * pixel (y + py - 1 + a, x + px - 1 + c)).  gn_stats (optional) covers the high-res output.  No residual / fp32 / geglu. */
  int32_t upsample2x;
  /* --- split-fp16 "precise" GEMM (near-tie re-scoring, see the block comment further down).  prec = 1: the A tensors and
   * w_ptr hold IEEE half.  Activations are [.., 2C] = hi plane | lo plane (a_channels = 2C); the K segments and Ktot
   * describe the LOGICAL K (channel ranges of the hi plane; the lo plane of the same channels lies C further right).  w_ptr is
   * K-BLOCK-MAJOR [Ktot/64][2 = Whi, Wlo][Npad][64], so that each operand tile of a K block is one contiguous run of HBM
   * (small-M launches stream every weight once).  Per K block the kernel stages the four tiles hi, lo, Whi, Wlo once and
   * issues hi x Whi + lo x Whi + hi x Wlo.  The accumulator is multiplied by acc_scale (weights are pre-scaled by a power of
   * two); `out` is fp32 (out_fp32) or split half [M, ld_out] with the lo plane at column out_lo_off; `residual` is a
   * split half tensor with its lo plane at column res_lo_off.  No gn_stats / geglu / upsample2x / strided sources. */
  int32_t prec;
  float acc_scale;
  int32_t out_lo_off, res_lo_off;
  /* prec only -- split-K: the K range of every output tile is cut into prec_splits slices (a function of the layer, never of
   * the batch: batch-size invariance) that run on different SMs and leave fp32 partial tiles in prec_partial
   * (>= prec_splits * ceil(M/128)*128 * Npad floats); a finishing kernel adds them in order.  prec_bn: N tile width
   * (256/192/128/64, 16 for Npad = 16; 0 = cost model). */
  int32_t prec_splits, prec_bn;
  float* prec_partial;
  /* optional: int32 [prec_ticket_len >= ceil(M/128) * Npad/64] arrival counters, zero-initialised ONCE (the kernel resets
   * them): the last K slice of an output tile to arrive adds all slices itself -- no finishing launch. */
  int32_t* prec_ticket;
  int32_t prec_ticket_len;
  /* epilogue activation applied to (acc + bias) before the residual / out_scale: 0 none, 1 quick_gelu x * sigmoid(1.702 x)
   * (CLIP's MLP, transformers activations.py QuickGELUActivation; 16-bit engine only) */
  int32_t act;
  /* optional: GroupNorm (affine, no activation) of the A operand inside the GEMM's operand path -- the `norm -> 1x1 conv`
   * pairs of the attention blocks (edm/training/networks.py:182-183 `qkv(norm2(x))`; diffusers `proj_in(norm(x))`) without the
   * normalised tensor being written: xf_mean_rstd fp32 [batch, xf_groups, 2] (b200ns_plan_add_gn_finalize), xf_gamma / xf_beta
   * fp32 [channels].  Needs ONE 1x1 segment over all channels of one unstrided source, 16-bit output, H*W = 64 or a multiple
   * of 128.  Bit-identical to b200ns_plan_add_gn_apply (silu = 0) followed by the plain GEMM. */
  const float* xf_mean_rstd;
  const float* xf_gamma;
  const float* xf_beta;
  int32_t xf_groups;
} b200ns_gemm_desc;
int b200ns_plan_add_gemm(b200ns_plan* p, const b200ns_gemm_desc* d);
/* Tuning aid: force the N tile width (64/128/192/256; 0 = cost model) of subsequently added bf16 GEMMs whose padded
 * width it divides.  Used by tools/profile_gemm_bn.py to measure the widths against each other. */
void b200ns_debug_force_tile_width(int bn);

/* GroupNorm statistics (networks.py:104-106): per-(sample, group) partial sums of x and x^2
 * over up to two concatenated sources, written as fp64 [batch, splits, groups, 2].
 * pre_add (fp32 [b_emb, C] or NULL) is added per channel first (networks.py:175). */
typedef struct {
  const void* x_ptr[2];
  int32_t x_channels[2];
  int32_t batch, HW, groups;
  const float* pre_add;
  int32_t ld_pre_add, b_emb;
  double* partial;           /* fp64 [batch, splits, groups, 2] */
  int32_t splits;
} b200ns_gn_stats_desc;
int b200ns_plan_add_gn_stats(b200ns_plan* p, const b200ns_gn_stats_desc* d);

/* GroupNorm statistics from the per-channel sums a producing GEMM left behind (b200ns_gemm_desc.gn_stats):
 * up to two concatenated sources; fp64 accumulation in a fixed order (bit-identical for identical inputs at
 * any batch position); pre_add as in gn_stats.  Writes fp32 [batch, groups, 2] = (mean, rstd)
 * (networks.py:104-106: var = E[x^2] - mean^2, biased, rstd = 1/sqrt(var + eps)). */
typedef struct {
  const float* stats_ptr[2]; /* fp32 [batch*HW/64, x_channels[i], 2] */
  int32_t x_channels[2];
  int32_t batch, HW, groups;
  const float* pre_add;
  int32_t ld_pre_add, b_emb;
  float eps;
  float* mean_rstd;          /* fp32 [batch, groups, 2] */
} b200ns_gn_finalize_desc;
int b200ns_plan_add_gn_finalize(b200ns_plan* p, const b200ns_gn_finalize_desc* d);

/* GroupNorm apply + optional FiLM + optional SiLU + optional 2x resample, bf16 -> bf16:
 *   y = act( (x+pre_add - mean)*rstd*gamma + beta ) ; FiLM: y = act( shift + norm*(scale+1) )
 * (networks.py:168, 173, 175, 182, 460).  resample: 0 none, 1 = 2x nearest up, 2 = 2x2 mean down
 * (networks.py:64-65, 82-85).  raw_out (optional) receives the resampled UN-normalised input
 * (the block's skip path, networks.py:159,178). */
typedef struct {
  const void* x_ptr[2];
  int32_t x_channels[2];
  int32_t batch, H, W, groups;
  const double* partial;     /* from b200ns_plan_add_gn_stats (or NULL with mean_rstd) */
  int32_t splits;
  float eps;
  const float* gamma;
  const float* beta;
  const float* pre_add;
  int32_t ld_pre_add;
  const float* film_scale;   /* fp32 [b_emb, ld_film] or NULL */
  const float* film_shift;
  int32_t ld_film, b_emb;
  int32_t silu;
  int32_t resample;
  void* out;                 /* bf16 [batch, H', W', C] */
  void* raw_out;             /* bf16 [batch, H', W', C] or NULL */
  const float* mean_rstd;    /* fp32 [batch, groups, 2] from gn_finalize; when set, `partial` is ignored */
  int32_t reverse;           /* walk direction, see b200ns_gemm_desc.reverse */
} b200ns_gn_apply_desc;
int b200ns_plan_add_gn_apply(b200ns_plan* p, const b200ns_gn_apply_desc* d);

/* b200ns_plan_add_gn_finalize + b200ns_plan_add_gn_apply as ONE launch for the low-resolution levels (H*W <= 256, where the
 * two kernels are latency-bound): a thread-block cluster of 8 CTAs per sample reduces the statistics (same code, same
 * order, same bits as the finalize kernel), exchanges (mean, rstd) through `mean_rstd` across a cluster barrier and
 * normalises the sample (GroupNorm.forward + the silu / addcmul glue of UNetBlock.forward, networks.py:104-106, 168-175).
 * `a->mean_rstd` must equal `f->mean_rstd`.  Bit-identical to the two separate ops. */
int b200ns_plan_add_gn_norm(b200ns_plan* p, const b200ns_gn_finalize_desc* f, const b200ns_gn_apply_desc* a);

/* Self-attention, head_dim 64 or 256 (networks.py:113-118, 182-185): for each (batch, head)
 * O = softmax(Q K^T / sqrt(64)) V with fp32 softmax; Q,K come from qk [batch*L, ld_qk]
 * (Q at column head*64, K at column k_col0 + head*64), V^T from vt [batch*heads, 64, L];
 * output bf16 [batch*L, ld_out] at column head*64. */
typedef struct {
  const void* qk;
  int32_t ld_qk, k_col0;
  const void* vt;            /* V^T [batch*heads, 64, L], or NULL: V row-major in `qk` at column v_col0 */
  void* out;
  int32_t ld_out;
  int32_t batch, heads, L;
  int32_t v_col0;
  int32_t head_dim;          /* 64 (default when 0) or 256 (one head, L <= 256: DDPM++, networks.py:263) */
  int32_t reverse;           /* walk direction, see b200ns_gemm_desc.reverse */
  /* --- SD-1.5 UNet (attention_processor.py AttnProcessor2_0): head_dim may be 128 / 192 = the true head dimension
   * (80 / 160; 40 -> 64) zero-padded by the projection weights; `scale` = true_head_dim^-0.5 (0 = head_dim^-0.5).
   * Cross-attention: K/V come from `kv` [kv_batch*kv_rows, ld_kv] (K at k_col0, V at v_col0, per head head*head_dim);
   * sample b attends context b / kv_div; each context has kv_rows rows (multiple of 128) of which the first
   * kv_len are tokens (the rest must be finite, e.g. zero: they are masked to -inf). */
  float scale;
  const void* kv;
  int32_t ld_kv, kv_batch, kv_rows, kv_len, kv_div;
} b200ns_attn_desc;
int b200ns_plan_add_attention(b200ns_plan* p, const b200ns_attn_desc* d);

/* fp32 linear layers of the embedding network (networks.py:39-43, 437-447, 170):
 *   out[r, n] = act( sum_k x[r,k]*W[n,k] + bias[n] + add[r,n] ), act: 0 none, 1 SiLU. */
typedef struct {
  const float* x;
  int32_t rows, K, ld_x;
  const float* w;            /* [N, K] */
  const float* bias;         /* [N] or NULL */
  const float* add;          /* [rows, ld_add] or NULL */
  int32_t ld_add;
  int32_t N;
  int32_t act;
  float* out;
  int32_t ld_out;
} b200ns_linear_desc;
int b200ns_plan_add_linear(b200ns_plan* p, const b200ns_linear_desc* d);

/* 3x3 im2col of the fp32 NCHW network input (Cin <= 7) into bf16 [batch*H*W, 64]
 * (K index = (kh*3+kw)*Cin + c, zero padded) feeding the first conv (networks.py:410). */
typedef struct {
  const float* x;
  void* out;
  int32_t batch, C, H, W;
} b200ns_im2col_desc;
int b200ns_plan_add_im2col(b200ns_plan* p, const b200ns_im2col_desc* d);

/* ------------------------------------------------------------------ precise (fp32-faithful) re-scoring path
 * The bf16 engine carries ~1e-4 of score noise; the reference evaluates the network in fp32 (networks.py:655-667) and the
 * north star demands the SAME selected indices (edm/main.py:842).  The contenders within delta of a round's best score are
 * therefore re-evaluated by these ops: activations in "split fp16" (v = hi + lo, two IEEE halves, NHWC [B,H,W,2C] with the
 * lo plane at channel offset C), GEMMs on the same tcgen05 main loop over 3 K segments (b200ns_gemm_desc.prec),
 * GroupNorm statistics in fp64 and SiLU / softmax with IEEE expf and division, attention in fp32 on the FMA pipe. */
typedef struct {
  const void* x_ptr[2];      /* split half NHWC [batch, H, W, 2*x_channels[i]]; second source optional (decoder concat) */
  int32_t x_channels[2];     /* logical channels C_i */
  int32_t batch, H, W, groups;
  float eps;
  const float* gamma;
  const float* beta;
  const float* pre_add;      /* fp32 [b_emb, ld_pre_add] or NULL (networks.py:175) */
  int32_t ld_pre_add;
  const float* film_scale;   /* fp32 [b_emb, ld_film] or NULL (networks.py:173) */
  const float* film_shift;
  int32_t ld_film, b_emb;
  int32_t silu, resample;    /* as b200ns_gn_apply_desc */
  void* out;                 /* split half [batch, H', W', 2C] */
  void* raw_out;             /* split half or NULL */
  float* mean_rstd;          /* fp32 [batch, groups, 2]: written by the stats op, read by the apply op */
  double* partial;           /* stats scratch: fp64 [batch, 64, groups, 2] */
  int32_t* ticket;           /* stats scratch: int32 [batch], zero-initialised once (the kernel resets it) */
} b200ns_gn_prec_desc;
/* per-(sample, group) mean / rstd in fp64 over hi + lo (+ pre_add) -> mean_rstd (networks.py:104-106) */
int b200ns_plan_add_gn_stats_prec(b200ns_plan* p, const b200ns_gn_prec_desc* d);
/* y = act(FiLM(norm(x))) with optional 2x resample, fp32 arithmetic, split half in / out */
int b200ns_plan_add_gn_apply_prec(b200ns_plan* p, const b200ns_gn_prec_desc* d);

/* Self-attention, head_dim 64, fp32 (networks.py:113-118): qkv split half [batch*L, ld] with the lo plane at column
 * lo_off; inside a plane Q at column head*64, K at k_col0 + head*64, V at v_col0 + head*64; out split half
 * [batch*L, ld_out] (head h at column h*64, lo plane at out_lo_off).  L % 64 == 0. */
typedef struct {
  const void* qkv;
  int32_t ld, lo_off, k_col0, v_col0;
  void* out;
  int32_t ld_out, out_lo_off;
  int32_t batch, heads, L;
  float scale;               /* 0 = 1/sqrt(64) */
} b200ns_attn_prec_desc;
int b200ns_plan_add_attention_prec(b200ns_plan* p, const b200ns_attn_prec_desc* d);

/* Experiment switch: 1 = the precise kernels write a zero lo plane (plain fp16 storage), 0 = split fp16 (default). */
int b200ns_debug_prec_nolo(int on);

/* 3x3 im2col of the fp32 NCHW network input (Cin*9 <= 64) into split half [batch*H*W, 128] = [64 hi taps | 64 lo taps]. */
int b200ns_plan_add_im2col_prec(b200ns_plan* p, const b200ns_im2col_desc* d);

/* ------------------------------------------------------------------ CLIP scorer glue
 * CLIPScorer (reference sd/scorers.py:149-213): `self.processor(images=...)` -> `self.clip.get_image_features` -> cosine
 * similarity with the prompt embedding.  The ViT tower runs on the plan ops above (GEMM with act = 1 for quick_gelu,
 * LayerNorm, attention with kv_len masking of the padded tokens); these are the remaining pieces. */
/* CLIPImageProcessor (PIL path) in integer arithmetic, bit-exact: Pillow's two-pass 8-bit bicubic resize (coefficients
 * scaled by 2^22, computed by the host: bounds int32 [S,2] = (first input index, taps), coeffs int32 [S, ks], already
 * restricted to the S centre-cropped output columns / rows), rescale + normalise through lut fp32 [3,256], written as the
 * patch matrix of the patch-embedding GEMM: patches (activation type) [batch*Lp, Kp], row b*Lp + 1 + py*(S/P) + px,
 * column c*P*P + i*P + j, columns >= 3*P*P zero.  Row b*Lp (class token) and rows > (S/P)^2 are NOT written (keep them 0).
 * tmp: uint8 [batch,3,H,S] workspace. */
typedef struct {
  const uint8_t* img;        /* [batch,3,H,W] */
  uint8_t* tmp;
  void* patches;
  const int32_t* h_bounds;
  const int32_t* h_coeffs;
  const int32_t* v_bounds;
  const int32_t* v_coeffs;
  const float* lut;
  int32_t batch, H, W, S, P, Lp, Kp, hks, vks;
} b200ns_clip_preprocess_desc;
int b200ns_plan_add_clip_preprocess(b200ns_plan* p, const b200ns_clip_preprocess_desc* d);
/* pooled = post_layernorm(last_hidden_state[:, 0]) (modeling_clip.py CLIPVisionTransformer): x activation rows
 * b*row_stride (elements), C channels -> out fp32 [batch, C], LayerNorm in fp32. */
int b200ns_plan_add_clip_pool_ln(b200ns_plan* p, const void* x, int64_t row_stride, const float* gamma, const float* beta,
                                 float* out, int32_t batch, int32_t C, float eps);
/* score[b] = sum_d (image_embeds[b,d]/|image_embeds[b]|) * (text_embeds[b or 0,d]/|text_embeds[..]|)   sd/scorers.py:178-213 */
int b200ns_plan_add_clip_cosine(b200ns_plan* p, const float* image_embeds, const float* text_embeds, int32_t text_rows,
                                float* score, int32_t batch, int32_t D);

/* ------------------------------------------------------------------ classifier scorer glue
 * ImageNetScorer (edm/scorers.py:143-174) around EncoderUNetModel (edm/unet.py:701-912): the torso runs
 * on the plan ops above; these are the remaining pieces. */
/* out = float(in) / 255                                               edm/scorers.py:153 */
int b200ns_plan_add_u8_to_f32(b200ns_plan* p, const uint8_t* in, float* out, int64_t n);
/* AttentionPool2d token construction (edm/unet.py:63-65): act bf16 [batch,T,C], pos fp32 [C,T+1] ->
 * tok bf16 [batch,T,C] (spatial tokens + pos) and tok0 fp32 [batch,C] (mean token + pos). */
int b200ns_plan_add_pool_tokens(b200ns_plan* p, const void* act, const float* pos, void* tok, float* tok0,
                                int32_t batch, int32_t T, int32_t C);
/* QKVAttention (edm/unet.py:388-407) for query token 0: qkv0 fp32 [batch,3C], kv bf16 [batch,T,2C] ->
 * out fp32 [batch,C]; heads of 64 channels. */
int b200ns_plan_add_pool_attention(b200ns_plan* p, const float* qkv0, const void* kv, float* out, int32_t batch,
                                   int32_t T, int32_t C);
/* scores[r] = softmax(logits[r,:K])[target[r]]                        edm/scorers.py:163-172 */
int b200ns_plan_add_softmax_gather(b200ns_plan* p, const float* logits, const int64_t* target, float* scores,
                                   int32_t rows, int32_t K);

/* ------------------------------------------------------------------ SD backend (BASELINE.json config 5)
 * Leaves of the SD-1.5-shaped UNet2DConditionModel that are not GEMM / GroupNorm / attention, as plan ops. */
/* nn.LayerNorm over the channels of a bf16 token tensor [rows, C] (attention.py BasicTransformerBlock norm1-3). */
int b200ns_plan_add_layernorm(b200ns_plan* p, const void* x, const float* gamma, const float* beta, void* out, int64_t rows,
                              int32_t C, float eps);
/* GEGLU (activations.py:117-123): in bf16 [rows, 2F] = [hidden | gate] -> out bf16 [rows, F] = hidden * gelu_erf(gate). */
int b200ns_plan_add_geglu(b200ns_plan* p, const void* in, void* out, int64_t rows, int32_t F);

/* --- SD VAE decoder (AutoencoderKL.decode, autoencoder_kl.py:287-320; vae.py Decoder) ---------------
 * Row softmax of the unfused mid-block attention (one head of dimension 512, attention_processor.py AttnProcessor2_0):
 * P[r, :] = softmax(scale * S[r, :]), fp32 [rows, L] -> bf16 [rows, L]. */
int b200ns_plan_add_softmax_rows(b200ns_plan* p, const float* S, void* P, int64_t rows, int32_t L, float scale);
/* post_quant_conv (1x1, C -> C, C <= 8) on fp32 NCHW latents [B, C, HW]; w [C, C], bias [C]. */
int b200ns_post_quant(const float* x, const float* w, const float* bias, float* out, int32_t B, int32_t C, int32_t HW,
                      void* stream);
/* Decoded image fp32 NHWC [B, HW, C] -> uint8 = trunc(clip(x*127.5+128, 0, 255)) (pipeline_stable_diffusion.py:1115):
 * integer channel sums chan_sums [B, 4] (feed b200ns_brightness_from_sums) and, if u8 != NULL, the uint8 NCHW image. */
int b200ns_image_sums(const float* img, uint32_t* chan_sums, uint8_t* u8, int64_t B, int32_t C, int32_t HW, void* stream);
/* F.interpolate(scale_factor=2, mode="nearest") on bf16 NHWC (upsampling.py Upsample2D.forward). */
int b200ns_plan_add_upsample2x(b200ns_plan* p, const void* in, void* out, int32_t batch, int32_t H, int32_t W, int32_t C);

/* Classifier-free guidance (pipeline_stable_diffusion.py:1073-1075) + the reference's DDIM step with supplied variance
 * noise (scheduling_ddim.py:398-460) for R candidate rows; row r descends from parent r / per_parent.
 *   eps = eu + g*(et-eu); x0 = (sample - sqrt(1-a_t) eps)/sqrt(a_t); prev = sqrt(a_prev) x0 + dir_coef eps + std noise
 * eps_u/eps_t: UNet output fp32 NHWC [P,H,W,C]; sample fp32 NCHW [P,C,H,W]; noise fp32 NCHW [R,...] or NULL.
 * prev: fp32 NCHW [R,...]; net_in (optional): the next UNet input, both CFG halves ([2R,...]). fp32, bit-exact vs torch. */
int b200ns_ddim_cfg_step(const float* eps_u, const float* eps_t, const float* sample, const float* noise, float* prev,
                         float* net_in, int64_t R, int32_t per_parent, int32_t C, int32_t HW, float guidance,
                         float sqrt_beta_t, float sqrt_alpha_t, float sqrt_alpha_prev, float dir_coef, float std_dev,
                         void* stream);
/* Guided eps of the second UNet call -> pred_original_sample of each candidate (scheduling_ddim.py:409) -> uint8
 * quantisation (pipeline...:1115) -> score = mean(u8/255) over (C,H,W) (sd/scorers.py:66-67, non-RGB branch).
 * pred_x0 (optional) fp32 NCHW [R,...]; sums (optional) int32 [R] = exact integer sum of the uint8 image. */
int b200ns_ddim_x0_score(const float* eps_u, const float* eps_t, const float* cand, float* pred_x0, int32_t* sums,
                         float* scores, int64_t R, int32_t C, int32_t HW, float guidance, float sqrt_beta_t,
                         float sqrt_alpha_t, void* stream);

/* Candidate noises of the SD eps_greedy / zero_order search (pipeline_stable_diffusion.py:1368-1379), fp32:
 * cand[n] = fresh[n] ? dirs[n] : pivot + ((dirs[n] / ||dirs[n]||_2 * u[n]) * lambda) * sqrt_e;  pivot [E],
 * dirs/cand [N, E], u [N] (the per-candidate torch.rand(1) scale draws), fresh uint8 [N] or NULL. */
int b200ns_sd_candidates(const float* pivot, const float* dirs, const float* u, const uint8_t* fresh, float* cand,
                         int64_t N, int64_t E, float lambda, float sqrt_e, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_NOISE_SEARCH_H */
