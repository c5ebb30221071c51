"""`torch.library` registration of the C-ABI entry points: the "thin C-ABI torch custom-op layer" of the north star.

Every sampler / scorer entry point of include/b200_noise_search.h is exposed as `torch.ops.b200ns.<name>` with a CUDA
implementation that forwards raw device pointers to libb200ns.so through `_lib.py` (ctypes) on torch's current stream.
There is deliberately NO CPU / CompositeImplicit implementation: calling an op on a CPU tensor raises
`NotImplementedError` from the dispatcher -- the product path has no fallback.

The U-Net engine is exposed as `torch.ops.b200ns.plan_run(int handle)`: a plan (all pointers, shapes and TMA
descriptors resolved at build time, captured as one CUDA graph) is identified by its integer handle.

`ops.py` routes its public functions through these ops (B200NS_TORCH_OPS=0 calls ctypes directly; the dispatcher adds
~2 us of host time per call, invisible next to a 30 ms GPU-bound search step).
"""
from __future__ import annotations

import torch

from . import _lib as L

_LIB = torch.library.Library('b200ns', 'DEF')
_DEFS = {
    # mutating ("out") forms: outputs are preallocated by the caller (ops.py), nothing is returned
    'heun_pre_': '(Tensor x_cur, Tensor eps, Tensor(a!) x_hat, Tensor(b!) net_in, float s, float c_in) -> ()',
    'heun_mid_': '(Tensor x_hat, Tensor F1, Tensor(a!) net_in2, Tensor(b!)? x_eul, float c_skip, float c_out, float t_hat, '
                 'float dt, float c_in_next) -> ()',
    'heun_post_': '(Tensor x_hat, Tensor F1, Tensor? F2, Tensor(a!)? x_next, Tensor(b!)? x0_u8, Tensor(c!)? chan_sums, '
                  'float c_skip1, float c_out1, float t_hat, float dt, float c_skip2, float c_out2, float t_next) -> ()',
    'quantize_u8_': '(Tensor x, Tensor(a!) out) -> ()',
    'channel_sums_u8_': '(Tensor img, Tensor(a!) sums) -> ()',
    'brightness_from_sums_': '(Tensor sums, Tensor(a!) scores, int C, int HW) -> ()',
    'argmax_first_': '(Tensor scores, int idx_base, Tensor(a!) idx, Tensor(b!)? key) -> ()',
    'gather_rows_': '(Tensor src, Tensor idx, Tensor(a!) dst) -> ()',
    'direction_norms_': '(Tensor dirs, Tensor(a!) norms) -> ()',
    'make_candidates_': '(Tensor pivot, Tensor dirs, Tensor norms, Tensor scale, Tensor? fresh_mask, Tensor? fresh, '
                        'Tensor(a!) cand) -> ()',
    'plan_run': '(int handle) -> ()',
}
for _name, _schema in _DEFS.items():
    _LIB.define(_name + _schema)


def _p(t):
    return None if t is None else t.data_ptr()


def _heun_pre_(x_cur, eps, x_hat, net_in, s, c_in):
    R, b = eps.shape[0], x_cur.shape[0]
    fn = L.lib().b200ns_heun_pre_f32noise if eps.dtype == torch.float32 else L.lib().b200ns_heun_pre
    L.check(fn(_p(x_cur), _p(eps), _p(x_hat), _p(net_in), R, b, eps[0].numel(), float(s), float(c_in), L.cur_stream()), 'heun_pre')


def _heun_mid_(x_hat, F1, net_in2, x_eul, c_skip, c_out, t_hat, dt, c_in_next):
    R, Cc, H, W = x_hat.shape
    L.check(L.lib().b200ns_heun_mid(_p(x_hat), _p(F1), _p(net_in2), _p(x_eul), R, Cc, H * W, float(c_skip), float(c_out),
                                    float(t_hat), float(dt), float(c_in_next), L.cur_stream()), 'heun_mid')


def _heun_post_(x_hat, F1, F2, x_next, x0_u8, chan_sums, c_skip1, c_out1, t_hat, dt, c_skip2, c_out2, t_next):
    R, Cc, H, W = x_hat.shape
    L.check(L.lib().b200ns_heun_post(_p(x_hat), _p(F1), _p(F2), _p(x_next), _p(x0_u8), _p(chan_sums), R, Cc, H * W,
                                     float(c_skip1), float(c_out1), float(t_hat), float(dt), float(c_skip2), float(c_out2),
                                     float(t_next), L.cur_stream()), 'heun_post')


def _quantize_u8_(x, out):
    L.check(L.lib().b200ns_quantize_u8(_p(x), _p(out), x.numel(), L.cur_stream()), 'quantize_u8')


def _channel_sums_u8_(img, sums):
    L.check(L.lib().b200ns_channel_sums_u8(_p(img), _p(sums), img.shape[0], img.shape[1], img[0, 0].numel(), L.cur_stream()),
            'channel_sums_u8')


def _brightness_from_sums_(sums, scores, C, HW):
    L.check(L.lib().b200ns_brightness_from_sums(_p(sums), _p(scores), sums.shape[0], C, HW, L.cur_stream()), 'brightness')


def _argmax_first_(scores, idx_base, idx, key):
    N, b = scores.shape
    L.check(L.lib().b200ns_argmax_first(_p(scores), N, b, idx_base, _p(idx), _p(key), L.cur_stream()), 'argmax')


def _gather_rows_(src, idx, dst):
    L.check(L.lib().b200ns_gather_rows(_p(src), _p(idx), _p(dst), src.shape[0], src.shape[1], src[0, 0].numel(),
                                       L.cur_stream()), 'gather_rows')


def _direction_norms_(dirs, norms):
    L.check(L.lib().b200ns_direction_norms(_p(dirs), _p(norms), dirs.shape[0], dirs[0].numel(), L.cur_stream()), 'norms')


def _make_candidates_(pivot, dirs, norms, scale, fresh_mask, fresh, cand):
    L.check(L.lib().b200ns_make_candidates(_p(pivot), _p(dirs), _p(norms), _p(scale), _p(fresh_mask), _p(fresh), _p(cand),
                                           dirs.shape[0], pivot.shape[0], dirs[0].numel(), L.cur_stream()), 'make_candidates')


def _plan_run(handle):
    L.check(L.lib().b200ns_plan_run(handle, L.cur_stream()), 'plan_run')


for _name in _DEFS:
    # CUDA only: no CPU kernel is registered, so the dispatcher itself refuses CPU tensors
    _LIB.impl(_name, globals()['_' + _name], 'CUDA' if _name != 'plan_run' else 'CompositeExplicitAutograd')

OPS = torch.ops.b200ns
NAMES = tuple(_DEFS)
