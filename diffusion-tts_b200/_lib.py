"""ctypes binding of libb200ns.so (C ABI declared in include/b200_noise_search.h).

The library is built in-tree by `build.py` (nvcc, sm_100a).  There is NO fallback: if the
shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# 16-bit storage type of activations / weights: IEEE half by default (3.5x less candidate-dependent score noise than
# bfloat16 at the same tensor-core rate, and the reference's own GPU dtype); B200NS_ACT=bf16 selects the bfloat16 build.
ACT_BF16 = os.environ.get('B200NS_ACT', 'fp16').lower() in ('bf16', 'bfloat16')
LIB_PATH = os.path.join(_HERE, 'libb200ns_bf16.so' if ACT_BF16 else 'libb200ns.so')
try:
    import torch as _torch
    ACT_DTYPE = _torch.bfloat16 if ACT_BF16 else _torch.float16
except ImportError:                      # symbol checks without torch
    ACT_DTYPE = None

c_i32, c_i64, c_f32, c_f64, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p


class KSeg(C.Structure):
    _fields_ = [('src', c_i32), ('taps', c_i32), ('cstart', c_i32), ('cblocks', c_i32)]


class GemmDesc(C.Structure):
    _fields_ = [('a_ptr', c_vp * 3), ('a_channels', c_i32 * 3), ('n_seg', c_i32), ('seg', KSeg * 8),
                ('batch', c_i32), ('H', c_i32), ('W', c_i32), ('w_ptr', c_vp), ('N', c_i32), ('Npad', c_i32),
                ('Ktot', c_i32), ('bias', c_vp), ('residual', c_vp), ('ld_res', c_i32), ('out_scale', c_f32),
                ('out', c_vp), ('ld_out', c_i32), ('out_fp32', c_i32), ('gn_stats', c_vp), ('reverse', c_i32), ('a_stride', c_i32 * 3), ('geglu', c_i32), ('upsample2x', c_i32),
                ('prec', c_i32), ('acc_scale', c_f32), ('out_lo_off', c_i32), ('res_lo_off', c_i32),
                ('prec_splits', c_i32), ('prec_bn', c_i32), ('prec_partial', c_vp), ('prec_ticket', c_vp),
                ('prec_ticket_len', c_i32), ('act', c_i32),
                ('xf_mean_rstd', c_vp), ('xf_gamma', c_vp), ('xf_beta', c_vp), ('xf_groups', c_i32)]


class GnStatsDesc(C.Structure):
    _fields_ = [('x_ptr', c_vp * 2), ('x_channels', c_i32 * 2), ('batch', c_i32), ('HW', c_i32), ('groups', c_i32),
                ('pre_add', c_vp), ('ld_pre_add', c_i32), ('b_emb', c_i32), ('partial', c_vp), ('splits', c_i32)]


class GnApplyDesc(C.Structure):
    _fields_ = [('x_ptr', c_vp * 2), ('x_channels', c_i32 * 2), ('batch', c_i32), ('H', c_i32), ('W', c_i32),
                ('groups', c_i32), ('partial', c_vp), ('splits', c_i32), ('eps', c_f32), ('gamma', c_vp),
                ('beta', c_vp), ('pre_add', c_vp), ('ld_pre_add', c_i32), ('film_scale', c_vp), ('film_shift', c_vp),
                ('ld_film', c_i32), ('b_emb', c_i32), ('silu', c_i32), ('resample', c_i32), ('out', c_vp),
                ('raw_out', c_vp), ('mean_rstd', c_vp), ('reverse', c_i32)]


class GnPrecDesc(C.Structure):
    _fields_ = [('x_ptr', c_vp * 2), ('x_channels', c_i32 * 2), ('batch', c_i32), ('H', c_i32), ('W', c_i32),
                ('groups', c_i32), ('eps', c_f32), ('gamma', c_vp), ('beta', c_vp), ('pre_add', c_vp),
                ('ld_pre_add', c_i32), ('film_scale', c_vp), ('film_shift', c_vp), ('ld_film', c_i32), ('b_emb', c_i32),
                ('silu', c_i32), ('resample', c_i32), ('out', c_vp), ('raw_out', c_vp), ('mean_rstd', c_vp), ('partial', c_vp),
                ('ticket', c_vp)]


class ClipPreprocessDesc(C.Structure):
    _fields_ = [('img', c_vp), ('tmp', c_vp), ('patches', c_vp), ('h_bounds', c_vp), ('h_coeffs', c_vp), ('v_bounds', c_vp),
                ('v_coeffs', c_vp), ('lut', c_vp), ('batch', c_i32), ('H', c_i32), ('W', c_i32), ('S', c_i32), ('P', c_i32),
                ('Lp', c_i32), ('Kp', c_i32), ('hks', c_i32), ('vks', c_i32)]


class AttnPrecDesc(C.Structure):
    _fields_ = [('qkv', c_vp), ('ld', c_i32), ('lo_off', c_i32), ('k_col0', c_i32), ('v_col0', c_i32), ('out', c_vp),
                ('ld_out', c_i32), ('out_lo_off', c_i32), ('batch', c_i32), ('heads', c_i32), ('L', c_i32), ('scale', c_f32)]


class GnFinalizeDesc(C.Structure):
    _fields_ = [('stats_ptr', c_vp * 2), ('x_channels', c_i32 * 2), ('batch', c_i32), ('HW', c_i32), ('groups', c_i32),
                ('pre_add', c_vp), ('ld_pre_add', c_i32), ('b_emb', c_i32), ('eps', c_f32), ('mean_rstd', c_vp)]


class AttnDesc(C.Structure):
    _fields_ = [('qk', c_vp), ('ld_qk', c_i32), ('k_col0', c_i32), ('vt', c_vp), ('out', c_vp), ('ld_out', c_i32),
                ('batch', c_i32), ('heads', c_i32), ('L', c_i32), ('v_col0', c_i32), ('head_dim', c_i32), ('reverse', c_i32), ('scale', c_f32), ('kv', c_vp), ('ld_kv', c_i32),
                ('kv_batch', c_i32), ('kv_rows', c_i32), ('kv_len', c_i32), ('kv_div', c_i32)]


class LinearDesc(C.Structure):
    _fields_ = [('x', c_vp), ('rows', c_i32), ('K', c_i32), ('ld_x', c_i32), ('w', c_vp), ('bias', c_vp),
                ('add', c_vp), ('ld_add', c_i32), ('N', c_i32), ('act', c_i32), ('out', c_vp), ('ld_out', c_i32)]


class Im2colDesc(C.Structure):
    _fields_ = [('x', c_vp), ('out', c_vp), ('batch', c_i32), ('C', c_i32), ('H', c_i32), ('W', c_i32)]


# name -> (restype, argtypes); mirrors include/b200_noise_search.h one to one
SIGNATURES = {
    'b200ns_last_error': (C.c_char_p, []),
    'b200ns_device_ok': (C.c_int, [C.c_int]),
    'b200ns_act_is_fp16': (C.c_int, []),
    'b200ns_heun_pre': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f64, c_f32, c_vp]),
    'b200ns_heun_pre_f32noise': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f64, c_f32, c_vp]),
    'b200ns_heun_mid': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_f32, c_f32, c_f64, c_f64, c_f32, c_vp]),
    'b200ns_heun_post': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_f32, c_f32, c_f64, c_f64,
                                   c_f32, c_f32, c_f64, c_vp]),
    'b200ns_quantize_u8': (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    'b200ns_channel_sums_u8': (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp]),
    'b200ns_brightness_from_sums': (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp]),
    'b200ns_argmax_first': (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    'b200ns_gather_rows': (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]),
    'b200ns_direction_norms': (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp]),
    'b200ns_make_candidates': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]),
    'b200ns_jpeg_size': (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_f32, c_f32, c_vp, c_vp, c_vp]),
    'b200ns_jpeg_tables_bytes': (C.c_int, []),
    'b200ns_plan_create': (c_vp, []),
    'b200ns_plan_destroy': (None, [c_vp]),
    'b200ns_plan_size': (C.c_int, [c_vp]),
    'b200ns_plan_set_pdl': (C.c_int, [c_vp, C.c_int]),
    'b200ns_plan_gemm_cols': (C.c_int, [c_vp, C.c_int]),
    'b200ns_debug_force_tile_width': (None, [C.c_int]),
    'b200ns_plan_set_lane': (C.c_int, [c_vp, C.c_int]),
    'b200ns_plan_run': (C.c_int, [c_vp, c_vp]),
    'b200ns_plan_instantiate_graph': (C.c_int, [c_vp]),
    'b200ns_plan_run_range': (C.c_int, [c_vp, C.c_int, C.c_int, c_vp]),
    'b200ns_plan_add_gemm': (C.c_int, [c_vp, C.POINTER(GemmDesc)]),
    'b200ns_plan_add_gn_stats': (C.c_int, [c_vp, C.POINTER(GnStatsDesc)]),
    'b200ns_plan_add_gn_apply': (C.c_int, [c_vp, C.POINTER(GnApplyDesc)]),
    'b200ns_plan_add_gn_finalize': (C.c_int, [c_vp, C.POINTER(GnFinalizeDesc)]),
    'b200ns_plan_add_gn_norm': (C.c_int, [c_vp, C.POINTER(GnFinalizeDesc), C.POINTER(GnApplyDesc)]),
    'b200ns_plan_add_attention': (C.c_int, [c_vp, C.POINTER(AttnDesc)]),
    'b200ns_plan_add_linear': (C.c_int, [c_vp, C.POINTER(LinearDesc)]),
    'b200ns_plan_add_im2col': (C.c_int, [c_vp, C.POINTER(Im2colDesc)]),
    'b200ns_debug_prec_nolo': (C.c_int, [C.c_int]),
    'b200ns_plan_add_im2col_prec': (C.c_int, [c_vp, C.POINTER(Im2colDesc)]),
    'b200ns_plan_add_gn_stats_prec': (C.c_int, [c_vp, C.POINTER(GnPrecDesc)]),
    'b200ns_plan_add_gn_apply_prec': (C.c_int, [c_vp, C.POINTER(GnPrecDesc)]),
    'b200ns_plan_add_attention_prec': (C.c_int, [c_vp, C.POINTER(AttnPrecDesc)]),
    'b200ns_plan_add_u8_to_f32': (C.c_int, [c_vp, c_vp, c_vp, c_i64]),
    'b200ns_plan_add_pool_tokens': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32]),
    'b200ns_plan_add_pool_attention': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32]),
    'b200ns_plan_add_softmax_gather': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32]),
    'b200ns_plan_add_layernorm': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_f32]),
    'b200ns_plan_add_clip_preprocess': (C.c_int, [c_vp, c_vp]),
    'b200ns_plan_add_clip_pool_ln': (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, c_i32, c_f32]),
    'b200ns_plan_add_clip_cosine': (C.c_int, [c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32]),
    'b200ns_plan_add_geglu': (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32]),
    'b200ns_plan_add_upsample2x': (C.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32]),
    'b200ns_ddim_cfg_step': (C.c_int, [c_vp] * 6 + [c_i64, c_i32, c_i32, c_i32] + [c_f32] * 6 + [c_vp]),
    'b200ns_plan_add_softmax_rows': (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_f32]),
    'b200ns_post_quant': (C.c_int, [c_vp] * 4 + [c_i32, c_i32, c_i32, c_vp]),
    'b200ns_image_sums': (C.c_int, [c_vp] * 3 + [c_i64, c_i32, c_i32, c_vp]),
    'b200ns_sd_candidates': (C.c_int, [c_vp] * 5 + [c_i64, c_i64, c_f32, c_f32, c_vp]),
    'b200ns_ddim_x0_score': (C.c_int, [c_vp] * 6 + [c_i64, c_i32, c_i32] + [c_f32] * 3 + [c_vp]),
}

_lib = None


def lib():
    """Load libb200ns.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(the CUDA extension is mandatory; there is no CPU fallback)')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if bool(handle.b200ns_act_is_fp16()) == ACT_BF16:
            raise RuntimeError(f'{LIB_PATH} was built for the other 16-bit storage type (B200NS_ACT): rebuild')
        _lib = handle
    return _lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = lib().b200ns_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'libb200ns {what} failed (rc={rc}): {msg}')


def ptr(t):
    """Device pointer of a CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('libb200ns takes CUDA tensors only (no CPU fallback)')
    return t.data_ptr()


def cur_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
