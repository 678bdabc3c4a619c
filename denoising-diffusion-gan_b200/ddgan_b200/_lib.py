"""ctypes binding of libddgan_b200.so (the C ABI declared in include/ddgan_b200.h).

There is no CPU fallback: if the library is missing, importing any op raises.  PyTorch is used only for device
memory, streams and autograd plumbing; every call below passes raw device pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'libddgan_b200.so')

MAX_SRC = 3


class ConvSrc(C.Structure):
    _fields_ = [('x', C.c_void_p), ('scale', C.c_void_p), ('shift', C.c_void_p), ('C', C.c_int), ('pitch', C.c_int), ('ss_stride', C.c_int), ('act', C.c_int),
                ('ntaps', C.c_int), ('padded', C.c_int), ('tap_dr', C.c_int8 * 9), ('tap_ds', C.c_int8 * 9),
                ('planes', C.c_void_p), ('planes_C', C.c_int), ('planes_c0', C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [('src', ConvSrc * MAX_SRC), ('nsrc', C.c_int), ('wpack', C.c_void_p), ('kb', C.c_int), ('nt', C.c_int), ('N', C.c_int),
                ('Hout', C.c_int), ('Wout', C.c_int), ('Hp', C.c_int), ('Wp', C.c_int), ('Cout', C.c_int),
                ('bias', C.c_void_p), ('addvec', C.c_void_p), ('addvec_stride', C.c_int), ('res', C.c_void_p),
                ('out_scale', C.c_float), ('out_act', C.c_int), ('out', C.c_void_p), ('out_mode', C.c_int),
                ('out_C', C.c_int), ('stats', C.c_void_p), ('precision', C.c_int), ('msub', C.c_int),
                ('force_linear', C.c_int), ('debug_prof', C.c_void_p), ('batch_rows', C.c_int), ('zero_border', C.c_int), ('out_planes', C.c_void_p),
                ('splitk_ws', C.c_void_p), ('splitk_ws_bytes', C.c_long)]


class WgradDesc(C.Structure):
    _fields_ = [('x', C.c_void_p), ('dy', C.c_void_p), ('dw', C.c_void_p), ('xpitch', C.c_int), ('dypitch', C.c_int),
                ('N', C.c_int), ('Hp', C.c_int), ('Wp', C.c_int), ('Cout', C.c_int), ('dy_cpad', C.c_int),
                ('Cin_real', C.c_int), ('Cin_pad', C.c_int), ('ntaps', C.c_int), ('tap_dr', C.c_int8 * 9),
                ('tap_ds', C.c_int8 * 9), ('s_co', C.c_long), ('s_ci', C.c_long), ('s_tap', C.c_long), ('precision', C.c_int), ('gain', C.c_float), ('debug_prof', C.c_void_p)]


class AttnDesc(C.Structure):
    _fields_ = [('qkv', C.c_void_p), ('w3pack', C.c_void_p), ('bias', C.c_void_p), ('res', C.c_void_p), ('out', C.c_void_p),
                ('stats', C.c_void_p), ('N', C.c_int), ('H', C.c_int), ('W', C.c_int), ('C', C.c_int), ('out_scale', C.c_float),
                ('precision', C.c_int)]


class PackItem(C.Structure):
    _fields_ = [('w', C.c_void_p), ('out', C.c_void_p), ('s_co', C.c_long), ('s_ci', C.c_long), ('s_tap', C.c_long),
                ('chunk_begin', C.c_long), ('cout', C.c_int), ('cin_real', C.c_int), ('cin_pad', C.c_int), ('ntaps', C.c_int),
                ('flip_taps', C.c_int), ('kb', C.c_int), ('stage_offset', C.c_int), ('total_stages', C.c_int),
                ('precision', C.c_int), ('nt', C.c_int)]


MLP_MAX_LAYERS = 8


class MlpDesc(C.Structure):
    _fields_ = [('W', C.c_void_p * MLP_MAX_LAYERS), ('b', C.c_void_p * MLP_MAX_LAYERS), ('dims', C.c_int * (MLP_MAX_LAYERS + 1)),
                ('nlayers', C.c_int), ('pixel_norm', C.c_int), ('act', C.c_int)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_long, C.c_float

_SIGNATURES = {
    'ddg_version': ([], _I),
    'ddg_upfirdn2d': ([_P, _P, _P, _L] + [_I] * 12 + [_P], _I),
    'ddg_upfirdn2d_out_size': ([_I] * 6, _I),
    'ddg_fused_bias_act': ([_P, _P, _P, _P, _L, _I, _I, _I, _I, _F, _F, _P], _I),
    'ddg_channel_sum': ([_P, _P, _I, _I, _I, _P], _I),
    'ddg_upfirdn2d_lp': ([_P, _P, _P, _L] + [_I] * 9 + [_P], _I),
    'ddg_fused_bias_act_lp': ([_P, _P, _P, _P, _L, _I, _I, _I, _I, _F, _F, _I, _P], _I),
    'ddg_groupnorm_fwd': ([_P] * 6 + [_I] * 4 + [_F, _I, _I, _P], _I),
    'ddg_groupnorm_bwd': ([_P] * 9 + [_I] * 6 + [_P], _I),
    'ddg_timestep_embedding': ([_P, _P, _I, _I, _F, _P], _I),
    'ddg_linear': ([_P] * 4 + [_I] * 8 + [_P], _I),
    'ddg_mlp_rows': ([_P, _I, _P, _I, _I, _P, _P], _I),
    'ddg_q_sample_pairs': ([_P] * 10 + [_I, _L, _P], _I),
    'ddg_sample_posterior': ([_P] * 8 + [_I, _L, _P], _I),
    'ddg_nchw_to_pnhwc': ([_P, _I, _P, _I, _P, _I, _I, _I, _I, _F, _F, _P], _I),
    'ddg_pnhwc_to_nchw': ([_P, _P] + [_I] * 6 + [_P], _I),
    'ddg_gn_prepare': ([_P, _I, _P, _I, _P, _P, _I, _I, _P, _P, _I, _I, _I, _F, _P], _I),
    'ddg_gn_prepare_bwd': ([_P, _P, _I, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P], _I),
    'ddg_fir_pnhwc': ([_P, _P, _P, _I, _P] + [_I] * 6 + [_F, _P], _I),
    'ddg_minibatch_stddev': ([_P, _P] + [_I] * 6 + [_P], _I),
    'ddg_spatial_sum': ([_P, _P] + [_I] * 5 + [_P], _I),
    'ddg_conv_last_launch_info': ([_P, _P, _P, _P], _I),
    'ddg_conv_last_launch_tma': ([], _I),
    'ddg_conv_last_launch_ksplit': ([], _I),
    'ddg_zero_border': ([_P] + [_I] * 4 + [_P], _I),
    'ddg_set_pdl': ([_I], _I),
    'ddg_softmax_rows': ([_P, _P, _L, _I, _I, _I, _P], _I),
    'ddg_softmax_rows_bwd': ([_P, _P, _P, _L, _I, _I, _F, _P], _I),
    'ddg_conv_tile_n': ([_I, _L], _I),
    'ddg_conv_set_nt256': ([_I], _I),
    'ddg_conv_packed_bytes': ([_I, _I, _I, _I, _I], _L),
    'ddg_conv_pack_weights': ([_P, _P, _I, _I, _I, _I, _L, _L, _L, _I, _I, _I, _I, _I, _I, _I, _L, _P], _I),
    'ddg_conv_pack_chunks': ([_I] * 5, _L),
    'ddg_conv_pack_batch': ([_P, _I, _L, _P], _I),
    'ddg_channel_grads_splits': ([_I] * 4, _I),
    'ddg_channel_grads': ([_P, _P, _P] + [_I] * 5 + [_F, _I, _P], _I),
    'ddg_s2d_weights': ([_P, _P, _I, _I, _I, _I, _P], _I),
    'ddg_images_to_u8': ([_P, _P, _I, _I, _I, _I, _F, _F, _P], _I),
    'ddg_conv2d_fwd': ([C.POINTER(ConvDesc), _P], _I),
    'ddg_split_planes': ([_P, _P, _I, _I, _I, _I, _I, _P], _I),
    'ddg_planes_bytes': ([_I, _I, _I, _I], _L),
    'ddg_conv2d_wgrad': ([C.POINTER(WgradDesc), _P], _I),
    'ddg_attention_fwd': ([C.POINTER(AttnDesc), _P], _I),
    'ddg_affine_act_fwd': ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    'ddg_affine_act_bwd': ([_P] * 6 + [_I] * 5 + [_P], _I),
    'ddg_gn_bwd_coeffs': ([_P, _P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], _I),
    'ddg_gn_bwd_dx': ([_P] * 6 + [_I] * 5 + [_P], _I),
    'ddg_stats_fwd': ([_P, _P, _I, _I, _I, _I, _P], _I),
    'ddg_stats_bwd': ([_P, _P, _P, _I, _I, _I, _I, _P], _I),
    'ddg_grad_norm_sq': ([_P, _L, _P, _P], _I),
    'ddg_adam_ema_step': ([_P, _P, _P, _P, _P, _L, _P, _P, _F, _F, _F, _F, _F, _F, _F, _P], _I),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(or denoising-diffusion-gan_b200/csrc/build.sh). There is no CPU fallback.')
        l = C.CDLL(LIB_PATH)
        l.ddg_last_error.restype = C.c_char_p
        l.ddg_last_error.argtypes = []
        for name, (args, res) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = res
        _lib = l
    return _lib


def exported_symbols():
    return ['ddg_last_error'] + list(_SIGNATURES)


def check(rc: int, what: str = ''):
    if rc != 0:
        raise RuntimeError(f'libddgan_b200 {what} failed ({rc}): {lib().ddg_last_error().decode()}')


def ptr(t):
    """Device address of a tensor; ints (pre-computed addresses of channel-sliced views) pass through."""
    if t is None or isinstance(t, int):
        return t
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


LP_DTYPES = {torch.float16: 1, torch.bfloat16: 2}


def require_cuda_f32(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('ddgan_b200 ops run on CUDA tensors only (no CPU fallback)')
        if t.dtype != torch.float32:
            raise RuntimeError(f'ddgan_b200 ops are fp32 at the boundary, got {t.dtype}')
