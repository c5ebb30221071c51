// C ABI of libb200ns.so (see include/b200_noise_search.h): launch wrappers, TMA descriptor
// construction and the plan runner.  No torch types, no CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_noise_search.h"
#include "attention.cuh"
#include "classifier.cuh"
#include "clip.cuh"
#include "gemm_conv.cuh"
#include "groupnorm.cuh"
#include "jpeg.cuh"
#include "precise.cuh"
#include "sampler.cuh"
#include "sdops.cuh"
#include "vae.cuh"

using namespace b200;

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return static_cast<int>(e);
}
#define CK(call)                                   \
  do {                                             \
    int _rc = check_cuda((call), #call);           \
    if (_rc) return _rc;                           \
  } while (0)
#define CK_LAUNCH(name)                            \
  do {                                             \
    int _rc = check_cuda(cudaGetLastError(), name); \
    if (_rc) return _rc;                           \
  } while (0)

cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Launch with the programmatic-stream-serialization attribute (PDL): only for kernels that call pdl_wait() before
// their first global-memory access (common.cuh).  OFF by default, B200NS_PDL=1 turns it on: measured on B200
// (profiles/r01_pdl_ab.txt) the eps_greedy step is 32.5 ms with PDL against 32.2 ms without -- the GPU runs
// power-capped (sw_power_cap, ~1.7 GHz), so the idle gaps between kernels are paid back as clocks and hiding them gains
// nothing.
int g_pdl_override = -1;   // >= 0 while the ops of a plan with its own PDL setting are being launched / captured
int pdl_mode() {          // 0 off (default), 1 every plan kernel, 2 only the latency-bound GroupNorm kernels
  if (g_pdl_override >= 0) return g_pdl_override;
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200NS_PDL");
    v = (e != nullptr && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
  }
  return v;
}
bool pdl_enabled() { return pdl_mode() == 1; }
// 2-CTA clusters for the GEMM (B200NS_CL2=1): see gemm_conv.cuh (multicast weight tiles)
// Launch mode of the large (>= 2 waves) BN = 192 / 256 GEMMs: 2 (default) = CTA pairs with tcgen05 cta_group::2 (M = 256 MMAs; each
// SM stages and reads half of the weight tile), 1 = 2-CTA clusters that multicast the weight tile, 0 = single CTAs.  All three
// give bit-identical outputs (tools/check_cg2.py); measured per ADM-64 forward at batch 64: 12.55 / 12.47 / 12.33 ms of GEMM time
// for modes 0 / 1 / 2 (profiles/r02_gemm_launch_modes.txt) -- the step is power-capped, so halving the B operand traffic buys 2 %.
int cl2_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200NS_CL2");
    v = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }
  return v;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_cluster2(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_light(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_mode() != 0 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int grid_for(int64_t work_items, int threads, int max_blocks = 148 * 16) {
  int64_t b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor, `rank` dims (innermost first), 128B swizzle, zero OOB fill.
int make_tmap(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box,
              const uint32_t* elem_strides = nullptr) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
             rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return fail(buf);
  }
  return 0;
}

// bf16 [rows, cols] matrix with row pitch `ld` elements; box = [box_rows][box_cols], 128B swizzle.
// explicit byte strides (dims[0] is contiguous): strided views such as one phase of a 2x-upsampled NHWC tensor
int make_tmap_strided(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i < rank - 1) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled (strided) failed (%d) rank=%d", static_cast<int>(r), rank);
    return fail(buf);
  }
  return 0;
}

int make_tmap_2d_ld(CUtensorMap* tm, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols,
                    uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return fail("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstr[1] = {ld * 2};
  const cuuint32_t bx[2] = {box_cols, box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled (2d, ld) failed (%d) cols=%llu rows=%llu ld=%llu", static_cast<int>(r),
             (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld);
    return fail(buf);
  }
  return 0;
}

// ------------------------------------------------------------------ plan ops
enum OpKind { OP_GEMM, OP_GN_STATS, OP_GN_APPLY, OP_GN_FINALIZE, OP_GN_NORM, OP_ATTN, OP_LINEAR, OP_IM2COL, OP_U8F32, OP_POOL_TOKENS, OP_POOL_ATTN,
              OP_SOFTMAX_GATHER, OP_LAYERNORM, OP_GEGLU, OP_UPSAMPLE2X, OP_SOFTMAX_ROWS,
              OP_GN_STATS_PREC, OP_GN_APPLY_PREC, OP_ATTN_PREC, OP_IM2COL_PREC,
              OP_CLIP_PREPROCESS, OP_CLIP_POOL_LN, OP_CLIP_COSINE };

struct GemmOp {
  CUtensorMap tmA[3], tmB, tmO, tmR;
  GemmArgs args;
  int BN;
  int grid;
  int cl2;
  int prec;                  // split-fp16 precise GEMM (precise.cuh)
  int xf;                    // GroupNorm of the A operand in the operand path (gemm_conv.cuh: gemm_transform)
  GemmPrecArgs pargs;
};
struct GnPrecOp {
  GnPrecArgs args;
  int grid;
};
struct AttnPrecOp {
  AttnPrecArgs args;
  dim3 grid;
};
struct GnStatsOp {
  GnStatsArgs args;
  dim3 grid;
  int threads;
};
struct GnApplyOp {
  GnApplyArgs args;
  dim3 grid;
  int threads;
};
struct GnFinalizeOp {
  GnFinalizeArgs args;
  dim3 grid;
  int n_pairs;
  int wide;
};
struct GnNormOp {          // finalize + apply in one cluster launch (low-resolution levels)
  GnApplyArgs apply;
  GnFinalizeArgs fin;
  dim3 grid;
  int threads;
};
struct AttnOp {
  CUtensorMap tmQ, tmK, tmV;
  AttnArgs args;
  int KT;
  int vrow;
  int head_dim;
  int variant;
  int alias;
  dim3 grid;
};
struct LinearOp {
  LinearArgs args;
  int grid;
};
struct Im2colOp {
  b200ns_im2col_desc d;
  int grid;
};

struct MiscOp {          // the small classifier kernels: plain pointers + a few ints
  const void* p0;
  const void* p1;
  void* p2;
  void* p3;
  int64_t n;
  int i0, i1, i2;
  float f0;
};

struct Op {
  OpKind kind;
  union {
    MiscOp misc;
    GemmOp gemm;
    GnStatsOp gns;
    GnApplyOp gna;
    GnFinalizeOp gnf;
    GnNormOp gnn;
    AttnOp attn;
    LinearOp lin;
    Im2colOp i2c;
    GnPrecOp gnp;
    AttnPrecOp attnp;
    ClipPreArgs clip;
  };
  int lane;     // 0 = main stream; k > 0: parallel branch k of the captured graph (see b200ns_plan_set_lane)
  Op() { memset(this, 0, sizeof(*this)); }
};

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN>
int launch_gemm_cg2(const GemmOp& g, cudaStream_t st) {          // CTA pair, tcgen05 cta_group::2 (M = 256)
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(gemm_conv_kernel<BN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  CK(launch_cluster2(gemm_conv_kernel<BN, true, true>, dim3(g.grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, g.tmA[0], g.tmA[1],
                     g.tmA[2], g.tmB, g.tmO, g.tmR, g.args));
  CK_LAUNCH("gemm_conv_kernel<cg2>");
  return 0;
}
template <int BN>
int launch_gemm_cl2(const GemmOp& g, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(gemm_conv_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  CK(launch_cluster2(gemm_conv_kernel<BN, true>, dim3(g.grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, g.tmA[0], g.tmA[1], g.tmA[2],
                     g.tmB, g.tmO, g.tmR, g.args));
  CK_LAUNCH("gemm_conv_kernel<cl2>");
  return 0;
}
template <int BN>
int launch_gemm_xf(const GemmOp& g, cudaStream_t st) {         // + 4 transform warps (GroupNorm of A in the operand path)
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(gemm_conv_kernel<BN, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  launch_pdl(gemm_conv_kernel<BN, false, false, true>, dim3(g.grid), dim3(Cfg::THREADS + 128), Cfg::SMEM_BYTES, st, g.tmA[0], g.tmA[1],
             g.tmA[2], g.tmB, g.tmO, g.tmR, g.args);
  CK_LAUNCH("gemm_conv_kernel<xf>");
  return 0;
}
template <int BN>
int launch_gemm_t(const GemmOp& g, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  if (g.xf) return launch_gemm_xf<BN>(g, st);
  if constexpr (BN == 192 || BN == 256) {
    if (g.cl2 == 2) return launch_gemm_cg2<BN>(g, st);
    if (g.cl2) return launch_gemm_cl2<BN>(g, st);
  }
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(gemm_conv_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  launch_pdl(gemm_conv_kernel<BN>, dim3(g.grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, g.tmA[0], g.tmA[1], g.tmA[2], g.tmB,
             g.tmO, g.tmR, g.args);
  CK_LAUNCH("gemm_conv_kernel");
  return 0;
}
int launch_gemm_small_n(const GemmOp& g, cudaStream_t st) {
  using Cfg = GemmCfg<16>;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(gemm_small_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  launch_pdl(gemm_small_n_kernel, dim3(g.grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, g.tmA[0], g.tmA[1], g.tmA[2], g.tmB, g.args);
  CK_LAUNCH("gemm_small_n_kernel");
  return 0;
}
template <int BN>
int launch_gemm_prec_t(const GemmOp& g, cudaStream_t st) {
  using Cfg = GemmPrecCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(gemm_prec_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  launch_pdl(gemm_prec_kernel<BN>, dim3(g.grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, g.tmA[0], g.tmA[1], g.tmA[2], g.tmB, g.args,
             g.pargs);
  CK_LAUNCH("gemm_prec_kernel");
  if (g.pargs.splits > 1) {          // add the K slices in order, then bias / residual / scale / split store
    const int64_t items = static_cast<int64_t>(g.args.M) * ((g.args.N + 7) / 8);
    launch_pdl(gemm_prec_finish_kernel, dim3(grid_for(items, 256, 148 * 8)), dim3(256), 0, st, g.args, g.pargs);
    CK_LAUNCH("gemm_prec_finish_kernel");
  }
  return 0;
}
int launch_gemm_prec(const GemmOp& g, cudaStream_t st) {
  switch (g.BN) {
    case 256: return launch_gemm_prec_t<256>(g, st);
    case 192: return launch_gemm_prec_t<192>(g, st);
    case 128: return launch_gemm_prec_t<128>(g, st);
    case 64: return launch_gemm_prec_t<64>(g, st);
    case 16: return launch_gemm_prec_t<16>(g, st);
  }
  return fail("bad BN (prec)");
}
int launch_gemm(const GemmOp& g, cudaStream_t st) {
  if (g.prec) return launch_gemm_prec(g, st);
  switch (g.BN) {
    case 256: return launch_gemm_t<256>(g, st);
    case 192: return launch_gemm_t<192>(g, st);
    case 128: return launch_gemm_t<128>(g, st);
    case 64: return launch_gemm_t<64>(g, st);
    case 16: return launch_gemm_small_n(g, st);
  }
  return fail("bad BN");
}

template <int KT, bool VROW>
int launch_attn_t(const AttnOp& o, cudaStream_t st) {
  using Cfg = AttnCfg<KT>;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(attention_kernel<KT, VROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  attention_kernel<KT, VROW><<<o.grid, 128, Cfg::SMEM_BYTES, st>>>(o.tmQ, o.tmK, o.tmV, o.args);
  CK_LAUNCH("attention_kernel");
  return 0;
}

// B200NS_ATTN_POLY=1 evaluates 1/4 of the softmax exponentials on the FMA pipe (ex2_poly) instead of the MUFU unit.
// Measured on B200 (profiles/r01_attention_poly_ab.txt): no gain (L=4096 D=64: 3.29 vs 3.29 ms; L=1024 D=128: 0.35 vs
// 0.32 ms) although ncu shows the XU pipe 70 % busy -- the softmax warps are issue/latency-bound, not MUFU-bound.  Off.
bool attn_poly() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200NS_ATTN_POLY");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
// B200NS_ATTN_ALIAS: 1 = head_dim-64 self-attention runs the aliased-P variant (KT = 64, 128 TMEM columns, 3 CTAs per SM)
bool attn_alias() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200NS_ATTN_ALIAS");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
template <int KT, int D, bool POLY, bool ALIAS = false>
int launch_attn_v3_p(const AttnOp& o, cudaStream_t st) {
  using Cfg = AttnCfg3<KT, D, ALIAS>;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(attention_kernel_v3<KT, D, POLY, ALIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  launch_pdl(attention_kernel_v3<KT, D, POLY, ALIAS>, o.grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, o.tmQ, o.tmK, o.tmV, o.args);
  CK_LAUNCH("attention_kernel_v3");
  return 0;
}
template <int KT, int D = 64>
int launch_attn_v3(const AttnOp& o, cudaStream_t st) {
  if constexpr (KT == 64 && D == 64) {
    if (o.alias) return launch_attn_v3_p<64, 64, false, true>(o, st);
  }
  return attn_poly() ? launch_attn_v3_p<KT, D, true>(o, st) : launch_attn_v3_p<KT, D, false>(o, st);
}

template <int KT>
int launch_attn_v2(const AttnOp& o, cudaStream_t st) {
  using Cfg = AttnCfg2<KT>;
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(attention_kernel_v2<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  attention_kernel_v2<KT><<<o.grid, 128, Cfg::SMEM_BYTES, st>>>(o.tmQ, o.tmK, o.tmV, o.args);
  CK_LAUNCH("attention_kernel_v2");
  return 0;
}

int run_op(const Op& op, cudaStream_t st) {
  switch (op.kind) {
    case OP_GEMM: return launch_gemm(op.gemm, st);
    case OP_GN_STATS:
      gn_stats_kernel<<<op.gns.grid, op.gns.threads, 0, st>>>(op.gns.args);
      CK_LAUNCH("gn_stats_kernel");
      return 0;
    case OP_GN_APPLY:
      if (op.gna.threads > 256)
        launch_pdl_light(gn_apply_kernel<true>, op.gna.grid, dim3(op.gna.threads), 0, st, op.gna.args);
      else
        launch_pdl_light(gn_apply_kernel<false>, op.gna.grid, dim3(op.gna.threads), 0, st, op.gna.args);
      CK_LAUNCH("gn_apply_kernel");
      return 0;
    case OP_GN_FINALIZE:
      if (op.gnf.wide)
        launch_pdl_light(gn_finalize_kernel<true>, dim3(op.gnf.n_pairs), dim3(256), 0, st, op.gnf.args, op.gnf.n_pairs);
      else
        launch_pdl_light(gn_finalize_kernel<false>, dim3(op.gnf.grid), dim3(256), 0, st, op.gnf.args, op.gnf.n_pairs);
      CK_LAUNCH("gn_finalize_kernel");
      return 0;
    case OP_GN_NORM: {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = op.gnn.grid;
      cfg.blockDim = dim3(op.gnn.threads);
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = GN_CLUSTER;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl_mode() != 0 ? 2 : 1;
      cudaLaunchKernelEx(&cfg, gn_norm_cluster_kernel, op.gnn.apply, op.gnn.fin);
      CK_LAUNCH("gn_norm_cluster_kernel");
      return 0;
    }
    case OP_ATTN:
      if (op.attn.head_dim == 256) {
        static bool attr_set = false;
        if (!attr_set) {
          CK(cudaFuncSetAttribute(attention_d256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg256::SMEM_BYTES));
          attr_set = true;
        }
        attention_d256_kernel<<<op.attn.grid, 128, AttnCfg256::SMEM_BYTES, st>>>(op.attn.tmQ, op.attn.tmK, op.attn.args);
        CK_LAUNCH("attention_d256_kernel");
        return 0;
      }
      if (op.attn.vrow && op.attn.variant == 2) return op.attn.KT == 128 ? launch_attn_v2<128>(op.attn, st) : launch_attn_v2<64>(op.attn, st);
      if (op.attn.vrow && op.attn.head_dim == 128) return launch_attn_v3<64, 128>(op.attn, st);
      if (op.attn.vrow && op.attn.head_dim == 192) return launch_attn_v3<64, 192>(op.attn, st);
      if (op.attn.vrow) return op.attn.KT == 128 ? launch_attn_v3<128>(op.attn, st) : launch_attn_v3<64>(op.attn, st);
      return op.attn.KT == 128 ? launch_attn_t<128, false>(op.attn, st) : launch_attn_t<64, false>(op.attn, st);
    case OP_GN_STATS_PREC:
      launch_pdl(gn_stats_prec_kernel, dim3(op.gnp.args.splits, op.gnp.args.batch), dim3((op.gnp.args.C / 8) * op.gnp.args.PY), 0, st,
                 op.gnp.args);
      CK_LAUNCH("gn_stats_prec_kernel");
      return 0;
    case OP_GN_APPLY_PREC:
      launch_pdl(gn_apply_prec_kernel, dim3(op.gnp.grid), dim3(256), 0, st, op.gnp.args);
      CK_LAUNCH("gn_apply_prec_kernel");
      return 0;
    case OP_ATTN_PREC: {
      static bool attr_set = false;
      if (!attr_set) {
        CK(cudaFuncSetAttribute(attention_prec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_PREC_SMEM));
        attr_set = true;
      }
      launch_pdl(attention_prec_kernel, op.attnp.grid, dim3(256), ATTN_PREC_SMEM, st, op.attnp.args);
      CK_LAUNCH("attention_prec_kernel");
      return 0;
    }
    case OP_IM2COL_PREC:
      im2col_c3_prec_kernel<<<op.i2c.grid, 256, 0, st>>>(op.i2c.d.x, reinterpret_cast<__half*>(op.i2c.d.out), op.i2c.d.batch,
                                                        op.i2c.d.C, op.i2c.d.H, op.i2c.d.W);
      CK_LAUNCH("im2col_c3_prec_kernel");
      return 0;
    case OP_LINEAR:
      launch_pdl(linear_kernel, dim3(op.lin.grid), dim3(256), 0, st, op.lin.args);
      CK_LAUNCH("linear_kernel");
      return 0;
    case OP_IM2COL:
      launch_pdl(im2col_c3_kernel, dim3(op.i2c.grid), dim3(256), 0, st, op.i2c.d.x, reinterpret_cast<act_t*>(op.i2c.d.out),
                 op.i2c.d.batch, op.i2c.d.C, op.i2c.d.H, op.i2c.d.W);
      CK_LAUNCH("im2col_c3_kernel");
      return 0;
    case OP_U8F32:
      u8_to_unit_f32_kernel<<<grid_for(op.misc.n, 256), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(op.misc.p0),
                                                                    reinterpret_cast<float*>(op.misc.p2), op.misc.n);
      CK_LAUNCH("u8_to_unit_f32_kernel");
      return 0;
    case OP_POOL_TOKENS:
      pool_tokens_kernel<<<op.misc.i0, 256, 0, st>>>(reinterpret_cast<const act_t*>(op.misc.p0),
                                                     reinterpret_cast<const float*>(op.misc.p1),
                                                     reinterpret_cast<act_t*>(op.misc.p2),
                                                     reinterpret_cast<float*>(op.misc.p3), op.misc.i1, op.misc.i2);
      CK_LAUNCH("pool_tokens_kernel");
      return 0;
    case OP_POOL_ATTN:
      pool_attention_kernel<<<dim3(op.misc.i2 / 64, op.misc.i0), 128, 0, st>>>(
          reinterpret_cast<const float*>(op.misc.p0), reinterpret_cast<const act_t*>(op.misc.p1),
          reinterpret_cast<float*>(op.misc.p2), op.misc.i1, op.misc.i2);
      CK_LAUNCH("pool_attention_kernel");
      return 0;
    case OP_SOFTMAX_GATHER:
      softmax_gather_kernel<<<op.misc.i0, 256, 0, st>>>(reinterpret_cast<const float*>(op.misc.p0),
                                                        reinterpret_cast<const int64_t*>(op.misc.p1),
                                                        reinterpret_cast<float*>(op.misc.p2), op.misc.i1);
      CK_LAUNCH("softmax_gather_kernel");
      return 0;
    case OP_LAYERNORM: {
      const act_t* lx = reinterpret_cast<const act_t*>(op.misc.p0);
      const float* lg = reinterpret_cast<const float*>(op.misc.p1);
      const float* lb = reinterpret_cast<const float*>(op.misc.p3);
      act_t* lo = reinterpret_cast<act_t*>(op.misc.p2);
      const int nchunk = op.misc.i0 / 8;
      const int64_t rows = op.misc.n;
      if (nchunk <= 40) {               // 4 rows per warp, 32 per CTA
        launch_pdl(layernorm_kernel<8, 5>, dim3(static_cast<unsigned>((rows + 31) / 32)), dim3(256), 0, st, lx, lg, lb, lo, rows, op.misc.i0, op.misc.f0);
      } else if (nchunk <= 80) {
        launch_pdl(layernorm_kernel<16, 5>, dim3(static_cast<unsigned>((rows + 15) / 16)), dim3(256), 0, st, lx, lg, lb, lo, rows, op.misc.i0, op.misc.f0);
      } else {
        launch_pdl(layernorm_kernel<32, 8>, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, st, lx, lg, lb, lo, rows, op.misc.i0, op.misc.f0);
      }
      CK_LAUNCH("layernorm_kernel");
      return 0;
    }
    case OP_GEGLU:
      launch_pdl(geglu_kernel, dim3(grid_for(op.misc.n * (op.misc.i0 / 8), 256, 148 * 32)), dim3(256), 0, st,
                 reinterpret_cast<const act_t*>(op.misc.p0), reinterpret_cast<act_t*>(op.misc.p2), op.misc.n, op.misc.i0);
      CK_LAUNCH("geglu_kernel");
      return 0;
    case OP_SOFTMAX_ROWS:
      softmax_rows_kernel<<<static_cast<unsigned>(op.misc.n), 256, 0, st>>>(reinterpret_cast<const float*>(op.misc.p0),
                                                                             reinterpret_cast<act_t*>(op.misc.p2),
                                                                             op.misc.i0, op.misc.f0);
      CK_LAUNCH("softmax_rows_kernel");
      return 0;
    case OP_CLIP_PREPROCESS: {
      const ClipPreArgs& c = op.clip;
      clip_resize_h_kernel<<<grid_for(static_cast<int64_t>(c.batch) * 3 * c.H * c.S, 256, 148 * 16), 256, 0, st>>>(c);
      CK_LAUNCH("clip_resize_h_kernel");
      clip_patches_kernel<<<grid_for(static_cast<int64_t>(c.batch) * c.G * c.G * c.Kp, 256, 148 * 16), 256, 0, st>>>(c);
      CK_LAUNCH("clip_patches_kernel");
      return 0;
    }
    case OP_CLIP_POOL_LN:
      clip_pool_ln_kernel<<<op.misc.i0, 256, 0, st>>>(reinterpret_cast<const act_t*>(op.misc.p0), op.misc.n,
                                                      reinterpret_cast<const float*>(op.misc.p1),
                                                      reinterpret_cast<const float*>(op.misc.p3),
                                                      reinterpret_cast<float*>(op.misc.p2), op.misc.i1, op.misc.f0);
      CK_LAUNCH("clip_pool_ln_kernel");
      return 0;
    case OP_CLIP_COSINE:
      clip_cosine_kernel<<<(op.misc.i0 + 3) / 4, 128, 0, st>>>(reinterpret_cast<const float*>(op.misc.p0),
                                                               reinterpret_cast<const float*>(op.misc.p1), op.misc.i2,
                                                               reinterpret_cast<float*>(op.misc.p2), op.misc.i0, op.misc.i1);
      CK_LAUNCH("clip_cosine_kernel");
      return 0;
    case OP_UPSAMPLE2X:
      launch_pdl(upsample2x_kernel, dim3(grid_for(static_cast<int64_t>(op.misc.i0) * op.misc.i1 * op.misc.i2 * 4 * (op.misc.n / 8), 256, 148 * 32)),
                 dim3(256), 0, st, reinterpret_cast<const act_t*>(op.misc.p0),
                 reinterpret_cast<act_t*>(op.misc.p2), op.misc.i0, op.misc.i1, op.misc.i2, static_cast<int>(op.misc.n));
      CK_LAUNCH("upsample2x_kernel");
      return 0;
  }
  return fail("bad op kind");
}

}  // namespace

struct b200ns_plan {
  std::vector<Op> ops;
  int cur_lane = 0;
  void push(Op& op) {
    op.lane = cur_lane;
    ops.push_back(op);
  }
  int pdl = -1;                             // -1: the process-wide B200NS_PDL setting; 0/1/2: this plan's own (small batches: 1)
  cudaGraphExec_t graph_exec = nullptr;     // optional: the whole plan captured once as a CUDA graph
  size_t graph_ops = 0;
};

extern "C" {

const char* b200ns_last_error(void) { return g_err.c_str(); }

int b200ns_act_is_fp16(void) {
#ifdef B200NS_ACT_BF16
  return 0;
#else
  return 1;
#endif
}

int b200ns_device_ok(int dev) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

// ------------------------------------------------------------------ sampler / scorer
int b200ns_heun_pre(const double* x_cur, const double* eps, double* x_hat, float* net_in, int64_t R, int64_t b,
                    int64_t E, double s, float c_in, void* stream) {
  if (E % 2) return fail("heun_pre: E must be even");
  const int64_t total = R * E;
  heun_pre_kernel<double><<<grid_for(total / 2, 256), 256, 0, S(stream)>>>(x_cur, eps, x_hat, net_in, total, b * E, s, c_in);
  CK_LAUNCH("heun_pre_kernel");
  return 0;
}

int b200ns_heun_pre_f32noise(const double* x_cur, const float* eps, double* x_hat, float* net_in, int64_t R, int64_t b,
                             int64_t E, double s, float c_in, void* stream) {
  if (E % 2) return fail("heun_pre: E must be even");
  const int64_t total = R * E;
  heun_pre_kernel<float><<<grid_for(total / 2, 256), 256, 0, S(stream)>>>(x_cur, eps, x_hat, net_in, total, b * E, s, c_in);
  CK_LAUNCH("heun_pre_kernel<float>");
  return 0;
}

int b200ns_heun_mid(const double* x_hat, const float* F1, float* net_in2, double* x_eul, int64_t R, int32_t C,
                    int32_t HW, float c_skip, float c_out, double t_hat, double dt, float c_in_next, void* stream) {
  HeunCoef k{};
  k.c_skip1 = c_skip;
  k.c_out1 = c_out;
  k.t_hat = t_hat;
  k.dt = dt;
  k.c_in_next = c_in_next;
  const int64_t total = R * C * HW;
  if (C == 3 && HW % 2 == 0) {          // RGB: two pixels x three channels per thread, vector loads / stores
    heun_mid_c3_kernel<<<grid_for(R * HW / 2, 256), 256, 0, S(stream)>>>(x_hat, F1, net_in2, x_eul, R * HW / 2, HW, k);
    CK_LAUNCH("heun_mid_c3_kernel");
    return 0;
  }
  heun_mid_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(x_hat, F1, net_in2, x_eul, total, C, HW, k);
  CK_LAUNCH("heun_mid_kernel");
  return 0;
}

int b200ns_heun_post(const double* x_hat, const float* F1, const float* F2, double* x_next, uint8_t* x0_u8,
                     uint32_t* chan_sums, int64_t R, int32_t C, int32_t HW, float c_skip1, float c_out1, double t_hat,
                     double dt, float c_skip2, float c_out2, double t_next, void* stream) {
  HeunCoef k{};
  k.c_skip1 = c_skip1;
  k.c_out1 = c_out1;
  k.t_hat = t_hat;
  k.dt = dt;
  k.c_skip2 = c_skip2;
  k.c_out2 = c_out2;
  k.t_next = t_next;
  if (R > 65535) return fail("heun_post: R > 65535");
  if (chan_sums != nullptr) CK(cudaMemsetAsync(chan_sums, 0, sizeof(uint32_t) * 4 * R, S(stream)));
  int chunks = static_cast<int>((2 * 148 + R - 1) / R);
  const int max_chunks = (HW + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (C == 3 && HW % 2 == 0) {
    const int max_c3 = (HW / 2 + 255) / 256;
    heun_post_c3_kernel<<<dim3(chunks > max_c3 ? max_c3 : chunks, static_cast<unsigned>(R)), 256, 0, S(stream)>>>(
        x_hat, F1, F2, x_next, x0_u8, chan_sums, HW, k);
    CK_LAUNCH("heun_post_c3_kernel");
    return 0;
  }
  heun_post_kernel<<<dim3(chunks, static_cast<unsigned>(R)), 256, 0, S(stream)>>>(x_hat, F1, F2, x_next, x0_u8, chan_sums,
                                                                                 C, HW, k);
  CK_LAUNCH("heun_post_kernel");
  return 0;
}

int b200ns_quantize_u8(const double* x, uint8_t* out, int64_t n, void* stream) {
  quantize_u8_kernel<<<grid_for(n, 256), 256, 0, S(stream)>>>(x, out, n);
  CK_LAUNCH("quantize_u8_kernel");
  return 0;
}

int b200ns_channel_sums_u8(const uint8_t* img, uint32_t* chan_sums, int64_t M, int32_t C, int32_t HW, void* stream) {
  if (M > 65535) return fail("channel_sums: M > 65535");
  if (C > 4) return fail("channel_sums: C > 4");
  CK(cudaMemsetAsync(chan_sums, 0, sizeof(uint32_t) * 4 * M, S(stream)));
  int chunks = static_cast<int>((2 * 148 + M - 1) / M);
  const int max_chunks = (HW + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  channel_sums_u8_kernel<<<dim3(chunks, static_cast<unsigned>(M)), 256, 0, S(stream)>>>(img, chan_sums, C, HW);
  CK_LAUNCH("channel_sums_u8_kernel");
  return 0;
}

int b200ns_brightness_from_sums(const uint32_t* chan_sums, float* scores, int64_t M, int32_t C, int32_t HW,
                                void* stream) {
  brightness_kernel<<<grid_for(M, 128), 128, 0, S(stream)>>>(chan_sums, scores, M, C, HW);
  CK_LAUNCH("brightness_kernel");
  return 0;
}

int b200ns_argmax_first(const float* scores, int64_t N, int64_t b, int64_t idx_base, int64_t* idx,
                        int64_t* packed_key, void* stream) {
  if (idx_base + N > 0xFFFFFFFFll) return fail("argmax_first: index overflow");
  argmax_first_kernel<<<static_cast<unsigned>(b), 32, 0, S(stream)>>>(scores, N, b, idx_base, idx, packed_key);
  CK_LAUNCH("argmax_first_kernel");
  return 0;
}

int b200ns_gather_rows(const double* src, const int64_t* idx, double* dst, int64_t N, int64_t b, int64_t E,
                       void* stream) {
  (void)N;
  gather_rows_kernel<<<dim3(grid_for(E, 256, 64), static_cast<unsigned>(b)), 256, 0, S(stream)>>>(src, idx, dst, b, E);
  CK_LAUNCH("gather_rows_kernel");
  return 0;
}

int b200ns_direction_norms(const double* dirs, double* norms, int64_t R, int64_t E, void* stream) {
  direction_norms_kernel<<<static_cast<unsigned>(R), 256, 0, S(stream)>>>(dirs, norms, E);
  CK_LAUNCH("direction_norms_kernel");
  return 0;
}

int b200ns_make_candidates(const double* pivot, const double* dirs, const double* norms, const float* scale,
                           const uint8_t* fresh_mask, const double* fresh, double* cand, int64_t R, int64_t b,
                           int64_t E, void* stream) {
  if (R > 65535) return fail("make_candidates: R > 65535");
  make_candidates_kernel<<<dim3(grid_for(E, 256, 16), static_cast<unsigned>(R)), 256, 0, S(stream)>>>(
      pivot, dirs, norms, scale, fresh_mask, fresh, cand, b, E);
  CK_LAUNCH("make_candidates_kernel");
  return 0;
}

int b200ns_jpeg_size(const uint8_t* img, const void* tables, int64_t M, int32_t H, int32_t W, float min_size,
                     float max_size, int32_t* sizes, float* scores, void* stream) {
  if (H % 16 || W % 16 || H > 64 || W > 64 || H <= 0 || W <= 0) return fail("jpeg_size: H, W must be multiples of 16, <= 64");
  const size_t smem = jpeg_smem_bytes(H, W);
  static size_t attr = 0;
  if (smem > attr) {
    CK(cudaFuncSetAttribute(jpeg_size_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  jpeg_size_kernel<<<static_cast<unsigned>(M), 128, smem, S(stream)>>>(img, reinterpret_cast<const JpegTables*>(tables), H, W,
                                                                       min_size, max_size, sizes, scores);
  CK_LAUNCH("jpeg_size_kernel");
  return 0;
}

int b200ns_jpeg_tables_bytes(void) { return static_cast<int>(sizeof(JpegTables)); }

// ------------------------------------------------------------------ plans
b200ns_plan* b200ns_plan_create(void) { return new b200ns_plan(); }
void b200ns_plan_destroy(b200ns_plan* p) {
  if (p != nullptr && p->graph_exec != nullptr) cudaGraphExecDestroy(p->graph_exec);
  delete p;
}

// Capture all ops of the plan into a CUDA graph (on a private stream; pointers and TMA descriptors are
// static), so that b200ns_plan_run costs one graph launch instead of one launch per kernel.
int b200ns_plan_instantiate_graph(b200ns_plan* p) {
  if (p->graph_exec != nullptr) {
    cudaGraphExecDestroy(p->graph_exec);
    p->graph_exec = nullptr;
  }
  // one eager pass first: sets the per-kernel attributes (max dynamic smem) outside the capture
  cudaStream_t cs;
  CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  for (size_t i = 0; i < p->ops.size(); ++i) {
    g_pdl_override = p->pdl;
    int rc = run_op(p->ops[i], cs);
    g_pdl_override = -1;
    if (rc) {
      cudaStreamDestroy(cs);
      return rc;
    }
  }
  CK(cudaStreamSynchronize(cs));
  // lanes > 0 become parallel branches: forked from the main stream at their first op, joined back before the
  // next main-lane op (and at the end).  Run eagerly (plan_run_range) the same ops simply execute in order.
  constexpr int MAX_LANES = 4;
  cudaStream_t ls[MAX_LANES] = {cs, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
  bool open[MAX_LANES] = {false, false, false, false};
  CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  for (int l = 1; l < MAX_LANES; ++l) {
    CK(cudaStreamCreateWithFlags(&ls[l], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev_join[l], cudaEventDisableTiming));
  }
  cudaGraph_t graph = nullptr;
  CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  int rc = 0;
  auto join_all = [&]() {
    for (int l = 1; l < MAX_LANES && rc == 0; ++l)
      if (open[l]) {
        rc = check_cuda(cudaEventRecord(ev_join[l], ls[l]), "cudaEventRecord(join)");
        if (rc == 0) rc = check_cuda(cudaStreamWaitEvent(cs, ev_join[l], 0), "cudaStreamWaitEvent(join)");
        open[l] = false;
      }
  };
  // B200NS_LANES_SEQ=1 (experiment): the sub-batches of a laned plan run one after the other on the main stream instead
  // of on parallel branches -- every kernel then works on 1/lanes of the batch, so a producer's output (<= 50 MB at
  // 64x64 with 2 lanes) is still in the 126 MB L2 when its consumer starts
  static const bool lanes_seq = [] {
    const char* e = getenv("B200NS_LANES_SEQ");
    return e != nullptr && e[0] == '1';
  }();
  for (size_t i = 0; i < p->ops.size() && rc == 0; ++i) {
    const int lane = lanes_seq ? 0 : p->ops[i].lane;
    if (lane <= 0 || lane >= MAX_LANES) {
      join_all();
      if (rc == 0) rc = run_op(p->ops[i], cs);
      continue;
    }
    if (!open[lane]) {
      rc = check_cuda(cudaEventRecord(ev_fork, cs), "cudaEventRecord(fork)");
      if (rc == 0) rc = check_cuda(cudaStreamWaitEvent(ls[lane], ev_fork, 0), "cudaStreamWaitEvent(fork)");
      open[lane] = true;
    }
    if (rc == 0) rc = run_op(p->ops[i], ls[lane]);
  }
  join_all();
  cudaError_t e = cudaStreamEndCapture(cs, &graph);
  if (rc == 0 && e != cudaSuccess) rc = check_cuda(e, "cudaStreamEndCapture");
  if (rc == 0) rc = check_cuda(cudaGraphInstantiate(&p->graph_exec, graph, 0), "cudaGraphInstantiate");
  if (graph != nullptr) cudaGraphDestroy(graph);
  for (int l = 1; l < MAX_LANES; ++l) {
    cudaStreamDestroy(ls[l]);
    cudaEventDestroy(ev_join[l]);
  }
  cudaEventDestroy(ev_fork);
  cudaStreamDestroy(cs);
  if (rc == 0) p->graph_ops = p->ops.size();
  return rc;
}
int b200ns_plan_set_lane(b200ns_plan* p, int lane) {
  if (lane < 0 || lane > 3) return fail("plan_set_lane: lane must be 0..3");
  p->cur_lane = lane;
  return 0;
}
int b200ns_plan_set_pdl(b200ns_plan* p, int mode) {
  if (mode < -1 || mode > 2) return fail("plan_set_pdl: mode must be -1 (process default), 0, 1 or 2");
  p->pdl = mode;
  return 0;
}
int b200ns_plan_size(const b200ns_plan* p) { return static_cast<int>(p->ops.size()); }
int b200ns_plan_gemm_cols(const b200ns_plan* p, int op) {
  if (op < 0 || op >= static_cast<int>(p->ops.size()) || p->ops[op].kind != OP_GEMM) return -1;
  return p->ops[op].gemm.args.N;
}

int b200ns_plan_run_range(b200ns_plan* p, int first, int last, void* stream) {
  if (first < 0 || last > static_cast<int>(p->ops.size()) || first > last) return fail("plan_run_range: bad range");
  g_pdl_override = p->pdl;
  for (int i = first; i < last; ++i) {
    int rc = run_op(p->ops[i], S(stream));
    if (rc) {
      g_pdl_override = -1;
      g_err = "plan op " + std::to_string(i) + ": " + g_err;
      return rc;
    }
  }
  g_pdl_override = -1;
  return 0;
}
int b200ns_plan_run(b200ns_plan* p, void* stream) {
  if (p->graph_exec != nullptr && p->graph_ops == p->ops.size()) {
    CK(cudaGraphLaunch(p->graph_exec, S(stream)));
    return 0;
  }
  return b200ns_plan_run_range(p, 0, static_cast<int>(p->ops.size()), stream);
}

int b200ns_debug_prec_nolo(int on) {
  const int v = on ? 1 : 0;
  CK(cudaMemcpyToSymbol(g_prec_nolo, &v, sizeof(int)));
  return 0;
}

static int g_force_bn = 0;
void b200ns_debug_force_tile_width(int bn) { g_force_bn = bn; }

// Relative cost of one K block of a 128 x c tile, MEASURED on B200 (tools/profile_gemm_bn.py, gpurun_out/gemm_bn.log:
// 3x3 convs at 64x64..8x8, batch 64): widths 192 and 256 run at the same rate per column (the tensor pipe needs ~2c
// cycles), 128 at ~0.78 of it and 64 at ~0.43 (the shared-memory operand feed, ~128 + c, and the per-tile epilogue).
static long gemm_per_kb(int c) { return c >= 192 ? 2 * c : (c == 128 ? 330 : 300); }

static int add_gemm_part(b200ns_plan* p, const b200ns_gemm_desc* d, int BN_forced, int ld_stats);
static int add_gemm_cols(b200ns_plan* p, const b200ns_gemm_desc* d);

// phase (py, px) of a fused "nearest 2x upsample + 3x3 conv" while its launches are being added (else phase < 0)
struct UpPhase {
  int phase = -1, dy0 = 0, dx0 = 0;
};
static UpPhase g_up;

int b200ns_plan_add_gemm(b200ns_plan* p, const b200ns_gemm_desc* d) {
  if (d->prec) {
    if (d->upsample2x || d->geglu || d->gn_stats != nullptr) return fail("gemm(prec): no upsample2x / geglu / gn_stats");
    if (!d->out_fp32 && (d->N % 32 || d->out_lo_off % 8)) return fail("gemm(prec): split output needs N % 32 == 0 and out_lo_off % 8 == 0");
    if (d->residual != nullptr && (d->res_lo_off % 8 || d->ld_res % 8)) return fail("gemm(prec): residual planes must be 16-byte aligned");
    if (d->prec_splits > 1 && d->prec_partial == nullptr) return fail("gemm(prec): split-K needs prec_partial");
    if (d->prec_ticket != nullptr && d->prec_ticket_len < ((d->batch * d->H * d->W + 127) / 128) * (d->Npad / 64))
      return fail("gemm(prec): prec_ticket too small (need ceil(M/128) * Npad/64 counters)");
    if (d->prec_bn != 0 && (d->Npad % d->prec_bn)) return fail("gemm(prec): prec_bn must divide Npad");
    int bn = d->prec_bn;
    if (bn == 0 && !d->out_fp32 && d->Npad % 64 == 0) {
      // tile width by the same waves x cost-per-K-block model as the bf16 GEMMs, over M tiles x N tiles x K slices (the
      // width may depend on the batch: it changes which SM computes an element, not the order of its K sum)
      const int m_tiles = (d->batch * d->H * d->W + 127) / 128;
      const int sp = d->prec_splits > 1 ? d->prec_splits : 1;
      const int cands[4] = {256, 192, 128, 64};
      long best = -1;
      for (int c : cands) {
        if (d->Npad % c) continue;
        const long items = static_cast<long>(m_tiles) * (d->Npad / c) * sp;
        const long cost = ((items + num_sms() - 1) / num_sms()) * gemm_per_kb(c);
        if (best < 0 || cost < best) best = cost, bn = c;
      }
    }
    return add_gemm_part(p, d, bn, d->N);
  }
  if (!d->upsample2x) return add_gemm_cols(p, d);
  // out [batch, 2H, 2W, ld_out] = conv3x3(nearest_up2(A)) as 4 phase launches of a 2x2-tap conv over the low-res A:
  // w_ptr = [4][Npad][Ktot], Ktot = 4 taps x channels of every segment, the 3x3 weights that land on the same source
  // pixel pre-summed (ops.pack_conv_up2); the desc's H, W are the LOW-res dims and its segments say taps = 9.
  if (d->out_fp32 || d->residual != nullptr || d->geglu) return fail("gemm(upsample2x): bf16 output, no residual, no geglu");
  if ((d->H * d->W) % 64) return fail("gemm(upsample2x): H*W must be a multiple of 64");
  for (int s = 0; s < d->n_seg; ++s)
    if (d->seg[s].taps != 9 || d->a_stride[d->seg[s].src] == 2) return fail("gemm(upsample2x): every segment must be a plain 3x3 conv");
  int rc = 0;
  for (int ph = 0; ph < 4 && rc == 0; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    b200ns_gemm_desc dp = *d;
    dp.w_ptr = static_cast<const char*>(d->w_ptr) + static_cast<size_t>(ph) * d->Npad * d->Ktot * 2;
    dp.out = static_cast<char*>(d->out) + (static_cast<size_t>(py) * 2 * d->W + px) * d->ld_out * 2;
    g_up.phase = ph;
    g_up.dy0 = py - 1;
    g_up.dx0 = px - 1;
    rc = add_gemm_cols(p, &dp);
  }
  g_up.phase = -1;
  return rc;
}

static int add_gemm_cols(b200ns_plan* p, const b200ns_gemm_desc* d) {
  // Tile width.  One launch uses one width c (a template parameter) and needs c | columns, so a weight matrix whose
  // padded width has no wide divisor (SD-1.5: 320 = 5 x 64) is covered by up to TWO launches over column slices
  // [0, n1*c1) and [n1*c1, Npad) of widths c1 != c2 (320 = 192 + 128, 640 = 2*192 + 256): pointers are shifted, the A operand is shared.
  // Cost model per part: (waves over the SMs) x (cycles per K block of one tile); an extra launch costs one more tail.
  if (d->out_fp32 || d->Npad % 64) return add_gemm_part(p, d, 0, d->N);
  if (d->geglu) {
    if (d->Npad % 128 || d->N % 128 || d->residual != nullptr || d->gn_stats != nullptr)
      return fail("gemm: geglu needs N and Npad multiples of 128, no residual and no gn_stats");
    return add_gemm_part(p, d, d->Npad % 256 == 0 ? 256 : 128, d->N);
  }
  if (g_force_bn > 0 && d->Npad % g_force_bn == 0) return add_gemm_part(p, d, g_force_bn, d->N);
  if (d->xf_mean_rstd != nullptr) {      // experiment: tile width of the GEMMs with GroupNorm in the operand path
    static int xf_bn = -1;
    if (xf_bn < 0) {
      const char* e = getenv("B200NS_XF_BN");
      xf_bn = e != nullptr ? atoi(e) : 0;
    }
    if (xf_bn > 0 && d->Npad % xf_bn == 0) return add_gemm_part(p, d, xf_bn, d->N);
  }
  const int m_tiles = (d->batch * d->H * d->W + 127) / 128;
  const int cands[4] = {256, 192, 128, 64};
  long best = -1;
  int bc1 = 0, bn1 = 0, bc2 = 0;
  auto waves = [&](long tiles) { return (tiles + num_sms() - 1) / num_sms(); };
  for (int c1 : cands) {
    for (int n1 = 1; n1 * c1 <= d->Npad; ++n1) {
      const int rest = d->Npad - n1 * c1;
      if (rest == 0) {
        const long cost = waves(static_cast<long>(m_tiles) * n1) * gemm_per_kb(c1);
        if (best < 0 || cost < best) best = cost, bc1 = c1, bn1 = n1, bc2 = 0;
        continue;
      }
      for (int c2 : cands) {
        if (c2 == c1 || rest % c2) continue;
        const long cost = waves(static_cast<long>(m_tiles) * n1) * gemm_per_kb(c1) +
                          waves(static_cast<long>(m_tiles) * (rest / c2)) * gemm_per_kb(c2) + gemm_per_kb(c2) / 2;
        if (best < 0 || cost < best) best = cost, bc1 = c1, bn1 = n1, bc2 = c2;
      }
    }
  }
  if (bc2 == 0) return add_gemm_part(p, d, bc1, d->N);
  const int split = bn1 * bc1;                      // first column of the second slice
  if (split >= d->N) {                              // the second slice would be padding only
    b200ns_gemm_desc d1 = *d;
    d1.Npad = split;
    return add_gemm_part(p, &d1, bc1, d->N);
  }
  b200ns_gemm_desc d1 = *d, d2 = *d;
  d1.N = split;
  d1.Npad = split;
  d2.N = d->N - split;
  d2.Npad = d->Npad - split;
  d2.w_ptr = static_cast<const char*>(d->w_ptr) + static_cast<size_t>(split) * d->Ktot * 2;
  if (d->bias) d2.bias = d->bias + split;
  if (d->residual) d2.residual = static_cast<const char*>(d->residual) + static_cast<size_t>(split) * 2;
  d2.out = static_cast<char*>(d->out) + static_cast<size_t>(split) * 2;
  if (d->gn_stats) d2.gn_stats = d->gn_stats + static_cast<size_t>(split) * 2;
  const int rc = add_gemm_part(p, &d1, bc1, d->N);
  if (rc) return rc;
  return add_gemm_part(p, &d2, bc2, d->N);
}

static int add_gemm_part(b200ns_plan* p, const b200ns_gemm_desc* d, int BN_forced, int ld_stats) {
  Op op;
  op.kind = OP_GEMM;
  GemmOp& g = op.gemm;
  GemmArgs& a = g.args;
  const int H = d->H, W = d->W;
  if (W <= 0 || (W <= 128 ? 128 % W : W % 128)) return fail("gemm: W must divide 128 or be a multiple of 128");
  const int x_chunks = W > 128 ? W / 128 : 1;      // rows wider than a tile: 128 consecutive pixels of one row per tile
  const int boxW = W > 128 ? 128 : W;
  int tileH = 128 / boxW;
  if (tileH > H) tileH = H;
  if (H % tileH) return fail("gemm: H not a multiple of the tile height");
  const int tileN = 128 / (boxW * tileH);
  a.x_chunks = x_chunks;
  a.H = H;
  a.W = W;
  a.tileH = tileH;
  a.tileN = tileN;
  a.tiles_per_img = (H * W >= 128) ? (H * W) / 128 : 0;
  a.M = d->batch * H * W;
  a.N = d->N;
  a.m_tiles = (a.M + 127) / 128;
  if (d->Npad % 16) return fail("gemm: Npad must be a multiple of 16");
  int BN = BN_forced;
  if (d->out_fp32) {
    BN = 16;
  } else if (!BN) {
    const int cands[4] = {256, 192, 128, 64};
    long best = -1;
    for (int c : cands) {
      if (d->Npad % c) continue;
      const long tiles = static_cast<long>(a.m_tiles) * (d->Npad / c);
      const long cost = ((tiles + num_sms() - 1) / num_sms()) * gemm_per_kb(c);
      if (best < 0 || cost < best) {
        best = cost;
        BN = c;
      }
    }
    if (!BN) BN = 16;
  }
  if (d->Npad % BN) return fail("gemm: no tile width divides Npad");
  g.BN = BN;
  a.n_tiles = d->Npad / BN;
  a.n_seg = d->n_seg;
  if (d->n_seg < 1 || d->n_seg > 8) return fail("gemm: n_seg out of range");
  int nkb = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    const b200ns_kseg& sg = d->seg[s];
    if (sg.taps != 1 && sg.taps != 9) return fail("gemm: taps must be 1 or 9");
    if (sg.src < 0 || sg.src > 2 || d->a_ptr[sg.src] == nullptr) return fail("gemm: bad segment source");
    // precise GEMM: segments address the hi plane, the lo plane of the same channels lies a_channels/2 further right
    if (sg.cstart % 64 || sg.cstart + sg.cblocks * 64 > (d->prec ? d->a_channels[sg.src] / 2 : d->a_channels[sg.src]))
      return fail("gemm: bad channel range");
    const int taps = g_up.phase >= 0 ? 4 : sg.taps;               // one phase of the fused upsample: 2x2 taps
    a.seg[s] = KSeg{sg.src, taps, sg.cstart, sg.cblocks, g_up.dy0, g_up.dx0};
    nkb += taps * sg.cblocks;
  }
  if (nkb * 64 != d->Ktot) return fail("gemm: Ktot does not match the K segments");
  a.nkb = nkb;
  a.bias = d->bias;
  a.fp16 = d->prec ? 1 : 0;
  g.prec = d->prec ? 1 : 0;
  g.pargs.acc_scale = d->prec ? d->acc_scale : 1.0f;
  g.pargs.out_lo_off = d->out_lo_off;
  g.pargs.res = d->prec ? reinterpret_cast<const __half*>(d->residual) : nullptr;
  g.pargs.ld_res = d->ld_res;
  g.pargs.res_lo_off = d->res_lo_off;
  {
    int sp = (d->prec && d->prec_splits > 1) ? d->prec_splits : 1;
    if (sp > nkb) sp = nkb;
    const int per = (nkb + sp - 1) / sp;
    sp = (nkb + per - 1) / per;                  // no empty slice
    g.pargs.splits = sp;
    g.pargs.kb_per_split = per;
    g.pargs.partial = d->prec ? d->prec_partial : nullptr;
    g.pargs.ld_partial = a.n_tiles * BN;
    g.pargs.ticket = (d->prec && sp > 1) ? d->prec_ticket : nullptr;
  }
  a.residual = d->prec ? nullptr : reinterpret_cast<const act_t*>(d->residual);
  a.ld_res = d->ld_res;
  a.out_scale = d->out_scale;
  a.out = d->out;
  a.ld_out = d->ld_out;
  a.out_fp32 = d->out_fp32;
  a.gn_stats = reinterpret_cast<float2*>(d->gn_stats);
  a.ld_stats = ld_stats;
  a.reverse = d->reverse;
  a.geglu = d->geglu;
  a.act = d->act;
  g.xf = 0;
  a.xf_mean_rstd = nullptr;
  if (d->xf_mean_rstd != nullptr) {
    const int HW = H * W;
    if (d->prec || d->upsample2x || d->out_fp32 || BN == 16) return fail("gemm(a_norm): 16-bit output GEMMs only");
    if (d->n_seg != 1 || d->seg[0].taps != 1 || d->seg[0].cstart != 0 || d->seg[0].cblocks * 64 != d->a_channels[d->seg[0].src] ||
        d->a_stride[d->seg[0].src] == 2)
      return fail("gemm(a_norm): a single 1x1 segment over all channels of one source");
    if (d->xf_gamma == nullptr || d->xf_beta == nullptr || d->xf_groups < 1 || d->a_channels[d->seg[0].src] % d->xf_groups)
      return fail("gemm(a_norm): gamma / beta / groups");
    if (HW < 64 || (HW < 128 && 128 % HW) || (HW >= 128 && HW % 128)) return fail("gemm(a_norm): H*W must be 64 or a multiple of 128");
    // the coefficient tables [2 tiles][2 samples][C] x (k_a, k_b) live in the shared memory of one operand stage
    if (32 * d->a_channels[d->seg[0].src] > 16384 + BN * 128) return fail("gemm(a_norm): too many channels for this tile width");
    g.xf = 1;
    a.xf_mean_rstd = reinterpret_cast<const float2*>(d->xf_mean_rstd);
    a.xf_gamma = d->xf_gamma;
    a.xf_beta = d->xf_beta;
    a.xf_groups = d->xf_groups;
    a.xf_cpg = d->a_channels[d->seg[0].src] / d->xf_groups;
    a.xf_batch = d->batch;
  }
  if (d->act != 0 && (d->act != 1 || d->geglu || d->upsample2x || d->prec)) return fail("gemm: act must be 0 or 1 (quick_gelu), without geglu / upsample2x / prec");
  if (d->gn_stats != nullptr && (BN == 16 || a.M % 64)) return fail("gemm: gn_stats needs bf16 output, N tiles >= 64 and M % 64 == 0");
  if (!d->out_fp32 && (d->ld_out % 8)) return fail("gemm: ld_out must be a multiple of 8 for bf16 output");
  if (d->residual != nullptr && (d->ld_res % 8)) return fail("gemm: ld_res must be a multiple of 8");
  a.out4d = 0;
  a.stats_in_rows = a.stats_img_rows = a.stats_off = 0;
  if (BN != 16 && !d->prec) {
    int rc;
    if (g_up.phase >= 0) {      // every other pixel of every other row of the [batch, 2H, 2W, ld_out] tensor (base = phase)
      const uint64_t ld = static_cast<uint64_t>(d->ld_out);
      const uint64_t dims[4] = {static_cast<uint64_t>(d->N), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                                static_cast<uint64_t>(d->batch)};
      const uint64_t strides[3] = {2 * ld * 2, 2 * (2 * static_cast<uint64_t>(W)) * ld * 2,
                                   (2 * static_cast<uint64_t>(H)) * (2 * static_cast<uint64_t>(W)) * ld * 2};
      const uint32_t box[4] = {64, static_cast<uint32_t>(boxW), static_cast<uint32_t>(tileH), static_cast<uint32_t>(tileN)};
      rc = make_tmap_strided(&g.tmO, d->out, 4, dims, strides, box);
      a.out4d = 1;
      a.stats_in_rows = (H * W) / 64;
      a.stats_img_rows = 4 * a.stats_in_rows;
      a.stats_off = g_up.phase * a.stats_in_rows;
    } else {
      rc = make_tmap_2d_ld(&g.tmO, d->out, static_cast<uint64_t>(d->geglu ? d->N / 2 : d->N), static_cast<uint64_t>(a.M),
                           static_cast<uint64_t>(d->ld_out), 64, 128);
    }
    if (rc) return rc;
    rc = make_tmap_2d_ld(&g.tmR, d->residual ? d->residual : d->out, static_cast<uint64_t>(d->N),
                         static_cast<uint64_t>(a.M), static_cast<uint64_t>(d->residual ? d->ld_res : d->ld_out), 64, 128);
    if (rc) return rc;
  }

  for (int i = 0; i < 3; ++i) {
    const bool used = d->a_ptr[i] != nullptr;
    const void* ptr = used ? d->a_ptr[i] : d->a_ptr[0];
    const int ch = used ? d->a_channels[i] : d->a_channels[0];
    const int sdn = (used ? d->a_stride[i] : d->a_stride[0]) == 2 ? 2 : 1;      // 2: this source is read with stride 2
    if (ch % 64) return fail("gemm: activation channels must be a multiple of 64");
    if (boxW * sdn > 256 || tileH * sdn > 256) return fail("gemm: strided box exceeds 256 elements");
    const uint64_t dims[4] = {static_cast<uint64_t>(ch), static_cast<uint64_t>(W) * sdn, static_cast<uint64_t>(H) * sdn,
                              static_cast<uint64_t>(d->batch)};
    const uint32_t box[4] = {64, static_cast<uint32_t>(boxW * sdn), static_cast<uint32_t>(tileH * sdn),
                             static_cast<uint32_t>(tileN)};
    const uint32_t es[4] = {1, static_cast<uint32_t>(sdn), static_cast<uint32_t>(sdn), 1};
    int rc = make_tmap(&g.tmA[i], ptr, 4, dims, box, es);
    if (rc) return rc;
    a.src_stride[i] = sdn;
  }
  // clusters of two CTAs on two M-adjacent tiles of one column slice: each fetches half of the weight tile for both
  g.cl2 = (cl2_enabled() && !d->prec && !g.xf && (BN == 192 || BN == 256) && a.m_tiles * a.n_tiles >= 2 * num_sms() && !d->out_fp32)
              ? cl2_enabled() : 0;
  if (d->prec) {        // K-block-major weights [nkb][hi, lo][Npad][64] (precise.cuh: gemm_prec_producer)
    for (int i = 0; i < 3; ++i) {
      if (d->a_ptr[i] != nullptr && d->a_channels[i] % 128) return fail("gemm(prec): split sources need 2 x (multiple of 64) channels");
      g.pargs.lo_off[i] = d->a_channels[i] / 2;
    }
    const uint64_t dims[2] = {64, 2 * static_cast<uint64_t>(nkb) * static_cast<uint64_t>(d->Npad)};
    const uint32_t box[2] = {64, static_cast<uint32_t>(BN)};
    int rc = make_tmap(&g.tmB, d->w_ptr, 2, dims, box);
    if (rc) return rc;
    g.pargs.n_pad = d->Npad;
  } else {
    const uint64_t dims[2] = {static_cast<uint64_t>(d->Ktot), static_cast<uint64_t>(d->Npad)};
    const uint32_t box[2] = {64, static_cast<uint32_t>(g.cl2 ? BN / 2 : BN)};
    int rc = make_tmap(&g.tmB, d->w_ptr, 2, dims, box);
    if (rc) return rc;
  }
  const int tiles = a.m_tiles * a.n_tiles * (g.prec ? g.pargs.splits : 1);
  g.grid = tiles < num_sms() ? tiles : num_sms();
  if (g.cl2) {
    const int super_tiles = ((a.m_tiles + 1) / 2) * a.n_tiles;
    const int max_pairs = num_sms() / 2;
    g.grid = 2 * (super_tiles < max_pairs ? super_tiles : max_pairs);
  }
  p->push(op);
  return 0;
}

int b200ns_plan_add_gn_stats(b200ns_plan* p, const b200ns_gn_stats_desc* d) {
  Op op;
  op.kind = OP_GN_STATS;
  GnStatsArgs& a = op.gns.args;
  a.x0 = reinterpret_cast<const act_t*>(d->x_ptr[0]);
  a.x1 = reinterpret_cast<const act_t*>(d->x_ptr[1]);
  a.C0 = d->x_channels[0];
  a.C1 = d->x_ptr[1] ? d->x_channels[1] : 0;
  a.C = a.C0 + a.C1;
  if (a.C % 8 || a.C0 % 8 || a.C > 2048) return fail("gn_stats: channels must be multiples of 8 and <= 2048");
  if (d->groups > 64 || a.C % d->groups) return fail("gn_stats: bad group count");
  a.HW = d->HW;
  a.groups = d->groups;
  a.cpg = a.C / d->groups;
  a.pre_add = d->pre_add;
  a.ld_pre_add = d->ld_pre_add;
  a.b_emb = d->b_emb > 0 ? d->b_emb : 1;
  a.partial = reinterpret_cast<double*>(d->partial);
  a.splits = d->splits;
  if (d->splits < 1 || d->HW % d->splits) return fail("gn_stats: splits must divide HW");
  const int VC = a.C / 8;
  a.PY = 256 / VC;
  if (a.PY < 1) a.PY = 1;
  op.gns.threads = VC * a.PY;
  op.gns.grid = dim3(d->splits, d->batch);
  p->push(op);
  return 0;
}

int b200ns_plan_add_gn_apply(b200ns_plan* p, const b200ns_gn_apply_desc* d) {
  Op op;
  op.kind = OP_GN_APPLY;
  GnApplyArgs& a = op.gna.args;
  a.x0 = reinterpret_cast<const act_t*>(d->x_ptr[0]);
  a.x1 = reinterpret_cast<const act_t*>(d->x_ptr[1]);
  a.C0 = d->x_channels[0];
  a.C1 = d->x_ptr[1] ? d->x_channels[1] : 0;
  a.C = a.C0 + a.C1;
  if (a.C % 8 || a.C0 % 8 || a.C > 4096) return fail("gn_apply: channels must be multiples of 8 and <= 4096");
  if (d->groups > 64 || a.C % d->groups) return fail("gn_apply: bad group count");
  if (d->resample == 2 && (d->H % 2 || d->W % 2)) return fail("gn_apply: odd size cannot be downsampled");
  a.H = d->H;
  a.W = d->W;
  a.groups = d->groups;
  a.cpg = a.C / d->groups;
  a.partial = reinterpret_cast<const double*>(d->partial);
  a.splits = d->splits;
  a.eps = d->eps;
  a.gamma = d->gamma;
  a.beta = d->beta;
  a.pre_add = d->pre_add;
  a.ld_pre_add = d->ld_pre_add;
  a.film_scale = d->film_scale;
  a.film_shift = d->film_shift;
  a.ld_film = d->ld_film;
  a.b_emb = d->b_emb > 0 ? d->b_emb : 1;
  a.silu = d->silu;
  a.resample = d->resample;
  a.out = reinterpret_cast<act_t*>(d->out);
  a.raw_out = reinterpret_cast<act_t*>(d->raw_out);
  a.mean_rstd = reinterpret_cast<const float2*>(d->mean_rstd);
  a.reverse = d->reverse;
  if (d->mean_rstd == nullptr && d->partial == nullptr) return fail("gn_apply: need partial or mean_rstd");
  const int VC = a.C / 8;
  a.PY = 256 / VC;
  if (a.PY < 1) a.PY = 1;
  op.gna.threads = VC * a.PY;
  const int dom = d->resample == 2 ? (d->H / 2) * (d->W / 2) : d->H * d->W;
  // pixels per thread: as many as keep >= ~4 CTAs per SM in flight (fewer prologues per byte), 4..32
  a.ITER = 32;
  while (a.ITER > 4 && static_cast<long long>((dom + a.PY * a.ITER - 1) / (a.PY * a.ITER)) * d->batch < 4LL * num_sms())
    a.ITER /= 2;
  const int per_cta = a.PY * a.ITER;
  op.gna.grid = dim3((dom + per_cta - 1) / per_cta, d->batch);
  p->push(op);
  return 0;
}

int b200ns_plan_add_gn_finalize(b200ns_plan* p, const b200ns_gn_finalize_desc* d) {
  Op op;
  op.kind = OP_GN_FINALIZE;
  GnFinalizeArgs& a = op.gnf.args;
  a.st0 = reinterpret_cast<const float2*>(d->stats_ptr[0]);
  a.st1 = reinterpret_cast<const float2*>(d->stats_ptr[1]);
  a.C0 = d->x_channels[0];
  a.C1 = d->stats_ptr[1] ? d->x_channels[1] : 0;
  a.C = a.C0 + a.C1;
  if (d->stats_ptr[0] == nullptr || d->mean_rstd == nullptr) return fail("gn_finalize: null pointer");
  if (d->groups < 1 || d->groups > 64 || a.C % d->groups) return fail("gn_finalize: bad group count");
  if (d->HW % 64) return fail("gn_finalize: HW must be a multiple of 64");
  a.HW = d->HW;
  a.groups = d->groups;
  a.cpg = a.C / d->groups;
  a.pre_add = d->pre_add;
  a.ld_pre_add = d->ld_pre_add;
  a.b_emb = d->b_emb > 0 ? d->b_emb : 1;
  a.eps = d->eps;
  a.mean_rstd = reinterpret_cast<float2*>(d->mean_rstd);
  op.gnf.n_pairs = d->groups * d->batch;
  op.gnf.grid = dim3((op.gnf.n_pairs + 7) / 8);
  op.gnf.wide = static_cast<long long>(a.HW / 64) * a.cpg > 4096 ? 1 : 0;      // a function of the shape only
  p->push(op);
  return 0;
}

// gn_finalize + gn_apply as ONE cluster launch (gn_norm_cluster_kernel): the two descriptors are exactly those of the two
// separate ops (apply.mean_rstd == finalize.mean_rstd), the result is bit-identical to running them back to back.
int b200ns_plan_add_gn_norm(b200ns_plan* p, const b200ns_gn_finalize_desc* f, const b200ns_gn_apply_desc* d) {
  b200ns_plan tmp;
  int rc = b200ns_plan_add_gn_finalize(&tmp, f);
  if (rc) return rc;
  rc = b200ns_plan_add_gn_apply(&tmp, d);
  if (rc) return rc;
  const Op& of = tmp.ops[0];
  const Op& oa = tmp.ops[1];
  if (of.gnf.wide || oa.gna.threads > 256) return fail("gn_norm: shape needs the wide kernels; use the separate ops");
  if (d->mean_rstd == nullptr || d->mean_rstd != f->mean_rstd) return fail("gn_norm: apply.mean_rstd must be finalize.mean_rstd");
  if (f->batch != d->batch || f->groups != d->groups || f->HW != d->H * d->W || of.gnf.args.C != oa.gna.args.C)
    return fail("gn_norm: the two descriptors disagree");
  Op op;
  op.kind = OP_GN_NORM;
  op.gnn.apply = oa.gna.args;
  op.gnn.fin = of.gnf.args;
  GnApplyArgs& a = op.gnn.apply;
  const int dom = d->resample == 2 ? (d->H / 2) * (d->W / 2) : d->H * d->W;
  a.ITER = (dom + GN_CLUSTER * a.PY - 1) / (GN_CLUSTER * a.PY);       // the sample's pixels over the cluster's CTAs
  op.gnn.threads = oa.gna.threads;
  op.gnn.grid = dim3(GN_CLUSTER, d->batch);
  p->push(op);
  return 0;
}

int b200ns_plan_add_attention(b200ns_plan* p, const b200ns_attn_desc* d) {
  Op op;
  op.kind = OP_ATTN;
  AttnOp& o = op.attn;
  if (d->L % 64) return fail("attention: L must be a multiple of 64");
  o.head_dim = d->head_dim > 0 ? d->head_dim : 64;
  o.KT = (d->L % 128 == 0) ? 128 : 64;
  o.args.out = reinterpret_cast<act_t*>(d->out);
  o.args.ld_out = d->ld_out;
  o.args.heads = d->heads;
  o.args.L = d->L;
  o.args.k_col0 = d->k_col0;
  o.args.v_col0 = d->v_col0;
  o.args.reverse = d->reverse;
  const uint64_t M = static_cast<uint64_t>(d->batch) * d->L;
  if (o.head_dim == 256) {
    if (d->heads != 1 || d->L > 256 || d->vt != nullptr) return fail("attention: head_dim 256 needs 1 head, L <= 256, row-major V");
    const uint64_t dims[2] = {static_cast<uint64_t>(d->ld_qk), M};
    const uint32_t boxq[2] = {64, 128};
    const uint32_t boxk[2] = {64, 64};
    int rc = make_tmap(&o.tmQ, d->qk, 2, dims, boxq);
    if (rc) return rc;
    rc = make_tmap(&o.tmK, d->qk, 2, dims, boxk);
    if (rc) return rc;
    o.vrow = 1;
    o.grid = dim3((d->L + 127) / 128, d->batch);
    p->push(op);
    return 0;
  }
  if (o.head_dim != 64 && o.head_dim != 128 && o.head_dim != 192)
    return fail("attention: (padded) head_dim must be 64, 128, 192 or 256");
  o.vrow = d->vt == nullptr ? 1 : 0;
  {
    const char* v = getenv("B200NS_ATTN");
    o.variant = (v != nullptr && v[0] == '2') ? 2 : 3;
  }
  const bool cross = d->kv != nullptr;
  if ((o.head_dim != 64 || cross || d->scale > 0.f) && (!o.vrow || o.variant != 3))
    return fail("attention: padded head dims / cross-attention / custom scale need the row-major-V v3 kernel");
  if (o.head_dim != 64) o.KT = 64;
  o.alias = (attn_alias() && o.head_dim == 64 && !cross && o.vrow && o.variant == 3 && d->L >= 256) ? 1 : 0;
  if (o.alias) o.KT = 64;
  if (cross) {
    if (d->kv_rows <= 0 || d->kv_rows % 128 || d->kv_len <= 0 || d->kv_len > d->kv_rows || d->kv_div <= 0 || d->kv_batch <= 0)
      return fail("attention: cross-attention needs kv_rows % 128 == 0, 0 < kv_len <= kv_rows, kv_div > 0, kv_batch > 0");
    o.KT = 64;
  }
  o.args.scale = d->scale > 0.f ? d->scale : 1.0f / sqrtf(static_cast<float>(o.head_dim));
  o.args.kv_rows = cross ? d->kv_rows : 0;
  o.args.kv_len = cross ? d->kv_len : 0;
  o.args.kv_div = cross ? d->kv_div : 1;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(d->ld_qk), M};
    const uint32_t boxq[2] = {64, 128};
    int rc = make_tmap(&o.tmQ, d->qk, 2, dims, boxq);
    if (rc) return rc;
  }
  {
    // K (and row-major V) tiles: [KT keys][64 d] boxes of the qkv matrix, or of the context K/V matrix (cross)
    const void* kvp = cross ? d->kv : d->qk;
    const uint64_t dims[2] = {static_cast<uint64_t>(cross ? d->ld_kv : d->ld_qk),
                              cross ? static_cast<uint64_t>(d->kv_batch) * d->kv_rows : M};
    const uint32_t boxk[2] = {64, static_cast<uint32_t>(o.KT)};
    int rc = make_tmap(&o.tmK, kvp, 2, dims, boxk);
    if (rc) return rc;
    if (o.vrow) {
      rc = make_tmap(&o.tmV, kvp, 2, dims, boxk);
      if (rc) return rc;
    }
  }
  if (!o.vrow) {
    const uint64_t dims[2] = {static_cast<uint64_t>(d->L), static_cast<uint64_t>(d->batch) * d->heads * 64};
    const uint32_t box[2] = {64, 64};
    int rc = make_tmap(&o.tmV, d->vt, 2, dims, box);
    if (rc) return rc;
  }
  o.grid = dim3((d->L + 127) / 128, d->batch * d->heads);
  p->push(op);
  return 0;
}

int b200ns_plan_add_linear(b200ns_plan* p, const b200ns_linear_desc* d) {
  Op op;
  op.kind = OP_LINEAR;
  LinearArgs& a = op.lin.args;
  a.x = d->x;
  a.rows = d->rows;
  a.K = d->K;
  a.ld_x = d->ld_x;
  a.w = d->w;
  a.bias = d->bias;
  a.add = d->add;
  a.ld_add = d->ld_add;
  a.N = d->N;
  a.act = d->act;
  a.out = d->out;
  a.ld_out = d->ld_out;
  const int64_t warps = static_cast<int64_t>(d->rows) * d->N;
  op.lin.grid = static_cast<int>((warps * 32 + 255) / 256);
  p->push(op);
  return 0;
}

int b200ns_plan_add_u8_to_f32(b200ns_plan* p, const uint8_t* in, float* out, int64_t n) {
  Op op;
  op.kind = OP_U8F32;
  op.misc.p0 = in;
  op.misc.p2 = out;
  op.misc.n = n;
  p->push(op);
  return 0;
}

int b200ns_plan_add_pool_tokens(b200ns_plan* p, const void* act, const float* pos, void* tok, float* tok0,
                                int32_t batch, int32_t T, int32_t C) {
  Op op;
  op.kind = OP_POOL_TOKENS;
  op.misc.p0 = act;
  op.misc.p1 = pos;
  op.misc.p2 = tok;
  op.misc.p3 = tok0;
  op.misc.i0 = batch;
  op.misc.i1 = T;
  op.misc.i2 = C;
  p->push(op);
  return 0;
}

int b200ns_plan_add_pool_attention(b200ns_plan* p, const float* qkv0, const void* kv, float* out, int32_t batch,
                                   int32_t T, int32_t C) {
  if (T > 127 || C % 64) return fail("pool_attention: T <= 127 and C % 64 == 0 required");
  Op op;
  op.kind = OP_POOL_ATTN;
  op.misc.p0 = qkv0;
  op.misc.p1 = kv;
  op.misc.p2 = out;
  op.misc.i0 = batch;
  op.misc.i1 = T;
  op.misc.i2 = C;
  p->push(op);
  return 0;
}

int b200ns_plan_add_softmax_gather(b200ns_plan* p, const float* logits, const int64_t* target, float* scores,
                                   int32_t rows, int32_t K) {
  Op op;
  op.kind = OP_SOFTMAX_GATHER;
  op.misc.p0 = logits;
  op.misc.p1 = target;
  op.misc.p2 = scores;
  op.misc.i0 = rows;
  op.misc.i1 = K;
  p->push(op);
  return 0;
}

int b200ns_plan_add_layernorm(b200ns_plan* p, const void* x, const float* gamma, const float* beta, void* out, int64_t rows,
                              int32_t C, float eps) {
  if (C % 8 || C > 2048) return fail("layernorm: C must be a multiple of 8 and <= 2048");
  Op op;
  op.kind = OP_LAYERNORM;
  op.misc.p0 = x;
  op.misc.p1 = gamma;
  op.misc.p3 = const_cast<float*>(beta);
  op.misc.p2 = out;
  op.misc.n = rows;
  op.misc.i0 = C;
  op.misc.f0 = eps;
  p->push(op);
  return 0;
}

int b200ns_plan_add_softmax_rows(b200ns_plan* p, const float* S, void* P, int64_t rows, int32_t L, float scale) {
  if (L % 8 || L > 8192 || rows <= 0) return fail("softmax_rows: L must be a multiple of 8 and <= 8192");
  Op op;
  op.kind = OP_SOFTMAX_ROWS;
  op.misc.p0 = S;
  op.misc.p2 = P;
  op.misc.n = rows;
  op.misc.i0 = L;
  op.misc.f0 = scale * 1.4426950408889634f;
  p->push(op);
  return 0;
}

int b200ns_post_quant(const float* x, const float* w, const float* bias, float* out, int32_t B, int32_t C, int32_t HW,
                      void* stream) {
  if (C < 1 || C > 8) return fail("post_quant: 1 <= C <= 8");
  post_quant_kernel<<<grid_for(static_cast<int64_t>(B) * HW, 256), 256, 0, S(stream)>>>(x, w, bias, out, B, C, HW);
  CK_LAUNCH("post_quant_kernel");
  return 0;
}

int b200ns_image_sums(const float* img, uint32_t* chan_sums, uint8_t* u8, int64_t B, int32_t C, int32_t HW, void* stream) {
  if (C < 1 || C > 4 || B > 65535) return fail("image_sums: 1 <= C <= 4, B <= 65535");
  CK(cudaMemsetAsync(chan_sums, 0, sizeof(uint32_t) * 4 * B, S(stream)));
  int chunks = static_cast<int>((4 * 148 + B - 1) / B);
  const int max_chunks = (HW + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  image_sums_kernel<<<dim3(chunks, static_cast<unsigned>(B)), 256, 0, S(stream)>>>(img, chan_sums, u8, C, HW);
  CK_LAUNCH("image_sums_kernel");
  return 0;
}

int b200ns_plan_add_geglu(b200ns_plan* p, const void* in, void* out, int64_t rows, int32_t F) {
  if (F % 8) return fail("geglu: F must be a multiple of 8");
  Op op;
  op.kind = OP_GEGLU;
  op.misc.p0 = in;
  op.misc.p2 = out;
  op.misc.n = rows;
  op.misc.i0 = F;
  p->push(op);
  return 0;
}

int b200ns_plan_add_upsample2x(b200ns_plan* p, const void* in, void* out, int32_t batch, int32_t H, int32_t W, int32_t C) {
  if (C % 8) return fail("upsample2x: C must be a multiple of 8");
  Op op;
  op.kind = OP_UPSAMPLE2X;
  op.misc.p0 = in;
  op.misc.p2 = out;
  op.misc.i0 = batch;
  op.misc.i1 = H;
  op.misc.i2 = W;
  op.misc.n = C;
  p->push(op);
  return 0;
}

int b200ns_ddim_cfg_step(const float* eps_u, const float* eps_t, const float* sample, const float* noise, float* prev,
                         float* net_in, int64_t R, int32_t per_parent, int32_t C, int32_t HW, float guidance,
                         float sqrt_beta_t, float sqrt_alpha_t, float sqrt_alpha_prev, float dir_coef, float std_dev,
                         void* stream) {
  if (R <= 0 || per_parent <= 0 || R % per_parent) return fail("ddim_cfg_step: R must be a positive multiple of per_parent");
  ddim_cfg_step_kernel<<<grid_for(R * C * HW, 256, 148 * 16), 256, 0, S(stream)>>>(
      eps_u, eps_t, sample, noise, prev, net_in, R, per_parent, C, HW, guidance, sqrt_beta_t, sqrt_alpha_t, sqrt_alpha_prev,
      dir_coef, std_dev);
  CK_LAUNCH("ddim_cfg_step_kernel");
  return 0;
}

int b200ns_ddim_x0_score(const float* eps_u, const float* eps_t, const float* cand, float* pred_x0, int32_t* sums,
                         float* scores, int64_t R, int32_t C, int32_t HW, float guidance, float sqrt_beta_t,
                         float sqrt_alpha_t, void* stream) {
  if (R <= 0) return fail("ddim_x0_score: R must be positive");
  ddim_x0_score_kernel<<<static_cast<unsigned>(R), 256, 0, S(stream)>>>(eps_u, eps_t, cand, pred_x0, sums, scores, C, HW,
                                                                        guidance, sqrt_beta_t, sqrt_alpha_t);
  CK_LAUNCH("ddim_x0_score_kernel");
  return 0;
}

int b200ns_sd_candidates(const float* pivot, const float* dirs, const float* u, const uint8_t* fresh, float* cand,
                         int64_t N, int64_t E, float lambda, float sqrt_e, void* stream) {
  if (N <= 0 || E <= 0) return fail("sd_candidates: N and E must be positive");
  sd_candidates_kernel<<<static_cast<unsigned>(N), 256, 0, S(stream)>>>(pivot, dirs, u, fresh, cand, E, lambda, sqrt_e);
  CK_LAUNCH("sd_candidates_kernel");
  return 0;
}

static int fill_gn_prec(GnPrecArgs& a, const b200ns_gn_prec_desc* d) {
  a.x0 = reinterpret_cast<const __half*>(d->x_ptr[0]);
  a.x1 = reinterpret_cast<const __half*>(d->x_ptr[1]);
  a.C0 = d->x_channels[0];
  a.C1 = d->x_ptr[1] ? d->x_channels[1] : 0;
  a.C = a.C0 + a.C1;
  if (a.x0 == nullptr || d->mean_rstd == nullptr) return fail("gn_prec: null pointer");
  if (a.C % 8 || a.C0 % 8) return fail("gn_prec: channels must be multiples of 8");
  if (d->groups < 1 || a.C % d->groups) return fail("gn_prec: bad group count");
  if (d->resample == 2 && (d->H % 2 || d->W % 2)) return fail("gn_prec: odd size cannot be downsampled");
  a.H = d->H;
  a.W = d->W;
  a.groups = d->groups;
  a.cpg = a.C / d->groups;
  a.eps = d->eps;
  a.gamma = d->gamma;
  a.beta = d->beta;
  a.pre_add = d->pre_add;
  a.ld_pre_add = d->ld_pre_add;
  a.film_scale = d->film_scale;
  a.film_shift = d->film_shift;
  a.ld_film = d->ld_film;
  a.b_emb = d->b_emb > 0 ? d->b_emb : 1;
  a.silu = d->silu;
  a.resample = d->resample;
  a.out = reinterpret_cast<__half*>(d->out);
  a.raw_out = reinterpret_cast<__half*>(d->raw_out);
  a.mean_rstd = reinterpret_cast<float2*>(d->mean_rstd);
  a.batch = d->batch;
  a.partial = d->partial;
  a.ticket = d->ticket;
  // pixel splits: a function of the image size ONLY (batch-size / batch-position invariance of the reduction order)
  const int HW = d->H * d->W;
  a.splits = HW / 16 < 1 ? 1 : (HW / 16 > 64 ? 64 : HW / 16);      // 16..64 pixels per CTA: short per-thread load chains
  if (HW % a.splits) a.splits = 1;
  const int VC = a.C / 8;
  if (a.C > 2048) return fail("gn_prec: more than 2048 channels");
  a.PY = 256 / VC < 1 ? 1 : 256 / VC;
  return 0;
}

int b200ns_plan_add_gn_stats_prec(b200ns_plan* p, const b200ns_gn_prec_desc* d) {
  Op op;
  op.kind = OP_GN_STATS_PREC;
  int rc = fill_gn_prec(op.gnp.args, d);
  if (rc) return rc;
  if (d->batch > 65535) return fail("gn_stats_prec: batch > 65535");
  if (d->partial == nullptr || d->ticket == nullptr) return fail("gn_stats_prec: partial / ticket scratch missing");
  p->push(op);
  return 0;
}

int b200ns_plan_add_gn_apply_prec(b200ns_plan* p, const b200ns_gn_prec_desc* d) {
  Op op;
  op.kind = OP_GN_APPLY_PREC;
  int rc = fill_gn_prec(op.gnp.args, d);
  if (rc) return rc;
  if (d->out == nullptr || d->gamma == nullptr || d->beta == nullptr) return fail("gn_apply_prec: null pointer");
  const GnPrecArgs& a = op.gnp.args;
  const int outH = a.resample == 1 ? a.H * 2 : (a.resample == 2 ? a.H / 2 : a.H);
  const int outW = a.resample == 1 ? a.W * 2 : (a.resample == 2 ? a.W / 2 : a.W);
  op.gnp.grid = grid_for(static_cast<int64_t>(a.batch) * outH * outW * (a.C / 8), 256, 148 * 8);
  p->push(op);
  return 0;
}

int b200ns_plan_add_attention_prec(b200ns_plan* p, const b200ns_attn_prec_desc* d) {
  if (d->L % 64 || d->L <= 0) return fail("attention_prec: L must be a positive multiple of 64");
  if (d->ld % 8 || d->lo_off % 8 || d->k_col0 % 8 || d->v_col0 % 8) return fail("attention_prec: columns must be 16-byte aligned");
  if (static_cast<int64_t>(d->batch) * d->heads > 65535) return fail("attention_prec: batch * heads > 65535");
  Op op;
  op.kind = OP_ATTN_PREC;
  AttnPrecArgs& a = op.attnp.args;
  a.qkv = reinterpret_cast<const __half*>(d->qkv);
  a.ld = d->ld;
  a.lo_off = d->lo_off;
  a.k_col0 = d->k_col0;
  a.v_col0 = d->v_col0;
  a.out = reinterpret_cast<__half*>(d->out);
  a.ld_out = d->ld_out;
  a.out_lo_off = d->out_lo_off;
  a.heads = d->heads;
  a.L = d->L;
  a.scale = d->scale > 0.f ? d->scale : 0.125f;
  op.attnp.grid = dim3(d->L / 64, d->batch * d->heads);
  p->push(op);
  return 0;
}

int b200ns_plan_add_im2col_prec(b200ns_plan* p, const b200ns_im2col_desc* d) {
  if (d->C * 9 > 64) return fail("im2col_prec: C*9 must be <= 64");
  Op op;
  op.kind = OP_IM2COL_PREC;
  op.i2c.d = *d;
  op.i2c.grid = grid_for(static_cast<int64_t>(d->batch) * d->H * d->W * 64, 256, 148 * 8);
  p->push(op);
  return 0;
}

int b200ns_plan_add_clip_preprocess(b200ns_plan* p, const b200ns_clip_preprocess_desc* d) {
  if (d->batch <= 0 || d->S <= 0 || d->P <= 0 || d->S % d->P) return fail("clip_preprocess: crop size must be a multiple of the patch size");
  const int G = d->S / d->P;
  if (d->Kp < 3 * d->P * d->P || d->Kp % 8 || d->Lp < G * G + 1) return fail("clip_preprocess: Kp >= 3*P*P (multiple of 8) and Lp >= G*G + 1 required");
  if (d->hks <= 0 || d->vks <= 0 || d->H <= 0 || d->W <= 0) return fail("clip_preprocess: bad sizes");
  Op op;
  op.kind = OP_CLIP_PREPROCESS;
  ClipPreArgs& c = op.clip;
  c.img = d->img;
  c.tmp = d->tmp;
  c.patches = reinterpret_cast<act_t*>(d->patches);
  c.hb = d->h_bounds;
  c.hk = d->h_coeffs;
  c.vb = d->v_bounds;
  c.vk = d->v_coeffs;
  c.lut = d->lut;
  c.batch = d->batch;
  c.H = d->H;
  c.W = d->W;
  c.S = d->S;
  c.P = d->P;
  c.G = G;
  c.Lp = d->Lp;
  c.Kp = d->Kp;
  c.hks = d->hks;
  c.vks = d->vks;
  p->push(op);
  return 0;
}

int b200ns_plan_add_clip_pool_ln(b200ns_plan* p, const void* x, int64_t row_stride, const float* gamma, const float* beta,
                                 float* out, int32_t batch, int32_t C, float eps) {
  if (batch <= 0 || C <= 0) return fail("clip_pool_ln: bad sizes");
  Op op;
  op.kind = OP_CLIP_POOL_LN;
  op.misc.p0 = x;
  op.misc.p1 = gamma;
  op.misc.p3 = const_cast<float*>(beta);
  op.misc.p2 = out;
  op.misc.n = row_stride;
  op.misc.i0 = batch;
  op.misc.i1 = C;
  op.misc.f0 = eps;
  p->push(op);
  return 0;
}

int b200ns_plan_add_clip_cosine(b200ns_plan* p, const float* image_embeds, const float* text_embeds, int32_t text_rows,
                                float* score, int32_t batch, int32_t D) {
  if (batch <= 0 || D <= 0 || (text_rows != 1 && text_rows != batch)) return fail("clip_cosine: text_rows must be 1 or batch");
  Op op;
  op.kind = OP_CLIP_COSINE;
  op.misc.p0 = image_embeds;
  op.misc.p1 = text_embeds;
  op.misc.p2 = score;
  op.misc.i0 = batch;
  op.misc.i1 = D;
  op.misc.i2 = text_rows;
  p->push(op);
  return 0;
}

int b200ns_plan_add_im2col(b200ns_plan* p, const b200ns_im2col_desc* d) {
  if (d->C * 9 > 64) return fail("im2col: C*9 must be <= 64");
  Op op;
  op.kind = OP_IM2COL;
  op.i2c.d = *d;
  op.i2c.grid = grid_for(static_cast<int64_t>(d->batch) * d->H * d->W * 8, 256);
  p->push(op);
  return 0;
}

}  // extern "C"
