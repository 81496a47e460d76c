"""ctypes binding of libtopaz_b200.so (the C ABI declared in include/topaz_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU fallback:
if the shared library is missing, or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtopaz_b200.so')
TPZ_TC_MAX_KB = 512


class TcKBlock(C.Structure):
    _fields_ = [('dx', C.c_int16), ('dy', C.c_int16), ('dz', C.c_int16), ('c0', C.c_int16), ('src', C.c_int32)]


class TpzTcSrc(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('N', C.c_int), ('D', C.c_int), ('H', C.c_int), ('W', C.c_int),
                ('C', C.c_int), ('ld', C.c_int), ('org', C.c_int * 3), ('kw', C.c_int), ('kh', C.c_int), ('lat', C.c_int), ('no_phase', C.c_int), ('lat_z', C.c_int)]


class TpzTcConvArgs(C.Structure):
    _fields_ = [
        ('nsrc', C.c_int), ('src', TpzTcSrc * 2), ('weights', C.c_void_p), ('KC', C.c_int), ('nkb', C.c_int),
        ('kb', TcKBlock * TPZ_TC_MAX_KB),
        ('N', C.c_int), ('Do', C.c_int), ('Ho', C.c_int), ('Wo', C.c_int), ('Co', C.c_int),
        ('TW', C.c_int), ('TH', C.c_int), ('lattice', C.c_int), ('phase_sel', C.c_int), ('lattice_z', C.c_int), ('phase_z', C.c_int),
        ('bias', C.c_void_p), ('neg_slope', C.c_float),
        ('res', C.c_void_p), ('res_scale', C.c_void_p),
        ('res_ld', C.c_int), ('res_D', C.c_int), ('res_H', C.c_int), ('res_W', C.c_int), ('res_org', C.c_int * 3),
        ('out', C.c_void_p), ('out_ld', C.c_int), ('out_coff', C.c_int),
        ('dot_w', C.c_void_p), ('dot_b', C.c_float), ('dot_out', C.c_void_p), ('dot_affine', C.c_void_p),
        ('oscale', C.c_void_p), ('range', C.c_void_p), ('out_lo', C.c_int),
    ]


class TpzLayerDesc(C.Structure):
    _fields_ = [('kind', C.c_int), ('cin', C.c_int), ('cout', C.c_int), ('k', C.c_int), ('dil0', C.c_int), ('dil1', C.c_int),
                ('slope0', C.c_float), ('slope1', C.c_float), ('w0', C.c_void_p), ('b0', C.c_void_p), ('w1', C.c_void_p),
                ('b1', C.c_void_p), ('proj', C.c_void_p), ('bn0', C.c_void_p), ('bn1', C.c_void_p), ('eps0', C.c_float),
                ('eps1', C.c_float)]


class TpzConvDesc(C.Structure):
    _fields_ = [('w', C.c_void_p), ('b', C.c_void_p), ('cout', C.c_int), ('cin', C.c_int), ('k', C.c_int)]


TPZ_UNET_MAX_DEPTH = 8


class TpzUnetDesc(C.Structure):
    _fields_ = [('dims', C.c_int), ('depth', C.c_int), ('enc', TpzConvDesc * TPZ_UNET_MAX_DEPTH), ('dec_a', TpzConvDesc * TPZ_UNET_MAX_DEPTH),
                ('dec_b', TpzConvDesc * TPZ_UNET_MAX_DEPTH), ('last', TpzConvDesc), ('slope', C.c_float), ('precision', C.c_int), ('host_weights', C.c_int)]


class TpzOpArgs(C.Structure):
    _fields_ = [('p', C.c_void_p * 6), ('n', C.c_longlong), ('i', C.c_int * 16), ('f', C.c_float * 4)]


LAUNCH_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)

_lib = None

_I, _F, _P, _LL, _D = C.c_int, C.c_float, C.c_void_p, C.c_longlong, C.c_double
_PROTOS = {
    'tpz_last_error': (C.c_char_p, []),
    'tpz_device_info': (_I, [C.POINTER(_I)] * 3),
    'tpz_tc_conv': (_I, [C.POINTER(TpzTcConvArgs), _P]),
    'tpz_tc_conv_v1': (_I, [C.POINTER(TpzTcConvArgs), _P]),
    'tpz_tc_conv_v2': (_I, [C.POINTER(TpzTcConvArgs), _P]),
    'tpz_model_create': (_I, [C.POINTER(TpzLayerDesc), _I, _P, _P, _I, C.POINTER(C.c_void_p), _P]),
    'tpz_model_update_weights': (_I, [_P, C.POINTER(TpzLayerDesc), _I, _P, _P, _P]),
    'tpz_model_destroy': (_I, [_P]),
    'tpz_model_timing': (_I, [_P, _I]),
    'tpz_model_timing_read': (_I, [_P, _P, _I, _P]),
    'tpz_model_step_buffers': (_I, [_P, _I, _P, _LL, _P, C.POINTER(_LL), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), _P]),
    'tpz_model_step_args': (_I, [_P, _I, C.POINTER(TpzTcConvArgs)]),
    'tpz_workspace_bytes': (_LL, [_P, _I, _I, _I]),
    'tpz_resnet_dense_forward': (_I, [_P, _P, _I, _I, _I, _P, _P, _LL, _P]),
    'tpz_unet_create': (_I, [C.POINTER(TpzUnetDesc), C.POINTER(C.c_void_p), _P]),
    'tpz_unet_destroy': (_I, [_P]),
    'tpz_unet_workspace_bytes': (_LL, [_P, _I, _I, _I, _I]),
    'tpz_unet_launch_count': (_I, [_P, _I, _I, _I, _I]),
    'tpz_unet2d_forward': (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _LL, _P]),
    'tpz_unet3d_forward': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _LL, _P]),
    'tpz_unet_set_launch_hook': (_I, [LAUNCH_HOOK, _P]),
    'tpz_unet_plan': (_I, [_P, _I, _I, _I, C.POINTER(TpzTcConvArgs), C.POINTER(_LL)]),
    'tpz_conv_first': (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P, _I, _P, _I, _P]),
    'tpz_range_scale': (_I, [_P, _LL, _P, _P, _P]),
    'tpz_im2col_first': (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P]),
    'tpz_conv_last': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _F, _I, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P]),
    'tpz_conv_generic': (_I, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F,
                              _P, _I, _I, _P, _I, _I, _I, _I, _P]),
    'tpz_maxpool2': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    'tpz_upsample_nearest': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    'tpz_meanstd': (_I, [_P, _LL, _I, _P, _P, _P]),
    'tpz_affine': (_I, [_P, _LL, _P, _I, _P, _P]),
    'tpz_sample_crops': (_I, [_I, C.c_ulonglong, C.c_ulonglong, _P, _P, _P, _P, _I, _P, _I, _P, _F, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    'tpz_make_crops': (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    'tpz_im2col3d_first': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P]),
    'tpz_conv_first_tc': (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _I, _F, _I, _P, _P, _P]),
    'tpz_conv_first_tc_supported': (_I, [_I, _I]),
    'tpz_gemm_f32': (_I, [_P, _LL, _I, _P, _I, _P, _P]),
    'tpz_gmm_sums': (_I, [_P, _LL, C.c_double, _P, _I, _P, _P]),
    'tpz_select_hist': (_I, [_P, _LL, _I, _P, _I, _P, _P]),
    'tpz_nms_flat': (_I, [_P, _LL, _P, _I, _F, _P, _P, _P, _I, C.POINTER(_I), _P]),
    'tpz_nms2d': (_I, [_P, _I, _I, _I, _F, _P, _P, _P, _I, C.POINTER(_I), _P]),
    'tpz_filter_f32': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _F, _P, _P]),
    'tpz_f32_to_f16': (_I, [_P, _LL, _P, _P]),
    'tpz_conv_fwd_f32': (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    'tpz_conv_dgrad_f32': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P]),
    'tpz_conv_wgrad_f32': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    'tpz_train_set_tf32': (_I, [_I]),
    'tpz_train_repack': (_I, [_P, _P, _I, _LL, _P, _P]),
    'tpz_conv_fwd_mma': (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    'tpz_conv_dgrad_mma': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P]),
    'tpz_conv_wgrad_mma': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'tpz_train_repack_tc': (_I, [_P, _P, _I, _LL, _P, _P]),
    'tpz_conv_fwd_tc': (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    'tpz_conv_dgrad_tc': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P]),
    'tpz_conv_wgrad_tc': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'tpz_conv_dgrad_tc_res': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _I, _P, _I, _I, _P]),
    'tpz_conv_wgrad_tc_bias': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    'tpz_first_fwd_f32': (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _I, _I, _P]),
    'tpz_first_wgrad_f32': (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    'tpz_first_fwd_tc': (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _I, _I, _P]),
    'tpz_first_wgrad_tc': (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    'tpz_bias_grad_f32': (_I, [_P, _LL, _I, _P, _P]),
    'tpz_cls_fwd_f32': (_I, [_P, _LL, _I, _P, _P, _P, _P]),
    'tpz_cls_bwd_f32': (_I, [_P, _LL, _I, _P, _P, _I, _P, _P, _P, _P]),
    'tpz_relu_bwd_f32': (_I, [_P, _P, _LL, _P]),
    'tpz_crop_add_f32': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P]),
    'tpz_bn_stats_f32': (_I, [_P, _LL, _I, _P, _P]),
    'tpz_bn_fwd_f32': (_I, [_P, _LL, _I, _P, _LL, _P, _P, _F, _F, _P, _P, _I, _P, _P, _P]),
    'tpz_bn_bwd_reduce_f32': (_I, [_P, _P, _LL, _I, _P, _P, _P]),
    'tpz_bn_bwd_f32': (_I, [_P, _P, _LL, _I, _P, _P, _LL, _P, _P, _P, _P, _P, _P]),
    'tpz_act_fwd_f32': (_I, [_P, _LL, _P, _F, _P, _P]),
    'tpz_act_bwd_f32': (_I, [_P, _P, _LL, _P, _F, _P, _P]),
    'tpz_dropout_fwd_f32': (_I, [_P, _LL, _F, C.c_ulonglong, C.c_ulonglong, _P, _P, _P]),
    'tpz_dropout_bwd_f32': (_I, [_P, _P, _LL, _F, _P]),
    'tpz_ge_binomial_loss_grad': (_I, [_P, _P, _I, _D, _D, _I, _I, _P, _P, _P]),
    'tpz_pu_objective_loss_grad': (_I, [_P, _P, _I, _I, _D, _D, _D, _D, _I, _I, _P, _P, _P]),
    'tpz_adam_step': (_I, [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _I, _F, _F, _P]),
    'tpz_adam_step_dev': (_I, [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _P, _F, _F, _P]),
}
EXPORTS = tuple(_PROTOS.keys())


def lib():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'topaz_b200: native library {LIB_PATH} is missing; run '
                               f'`python -c "import __graft_entry__ as g; g.build()"` (no CPU fallback exists)')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)       # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().tpz_last_error()
        raise RuntimeError(f'topaz_b200 native call failed ({rc}): {msg.decode() if msg else "?"}')
