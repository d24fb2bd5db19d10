"""ctypes binding of libynet_b200.so (the C ABI declared in include/ynet_b200.h).

There is NO CPU fallback: if the shared object is missing and cannot be built, importing the ops
fails loudly.  torch is used only for device memory and streams; the library sees raw pointers.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint64,
                    c_void_p)

from . import _build

_LIB = None

SRC_DIRECT, SRC_POOL2, SRC_UP2 = 0, 1, 2
MAX_SOURCES = 4


class ConvSrc(Structure):
    _fields_ = [('ptr', c_void_p), ('channels', c_int32), ('mode', c_int32), ('batch_stride', c_int64),
                ('batch_mod', c_int64)]


class TcSrc(Structure):
    _fields_ = [('ptr', c_void_p), ('channels_pad', c_int32), ('batch_mod', c_int32), ('batch_stride', c_int64),
                ('center_only', c_int32), ('chunks_stored', c_int32),
                ('padded', c_int32), ('tap_mask', c_int32)]


class TcWpSrc(Structure):
    _fields_ = [('tmpl_c8', c_void_p), ('coords', c_void_p), ('th', c_int32), ('tw', c_int32), ('n_ch', c_int32),
                ('level', c_int32)]


class YnetError(RuntimeError):
    pass


_P = c_void_p
_I = c_int32
_L = c_int64
_F = c_float

_PROTOS = {
    'ynet_version': (c_int, []),
    'ynet_last_error_string': (c_char_p, []),
    'ynet_device_info': (c_int, [POINTER(c_int32)] * 3),
    'ynet_rasterize_patches': (c_int, [_P, _I, _I, _P, _I, _P, _I, _I, _P, _P]),
    'ynet_rasterize_dist_analytic': (c_int, [_I, _P, _I, _P, _I, _I, _P]),
    'ynet_create_dist_template': (c_int, [_I, _P, _P]),
    'ynet_avgpool_pyramid': (c_int, [_P, _I, _I, _I, _I, POINTER(c_void_p), _P]),
    'ynet_softargmax2d_workspace_bytes': (_L, [_I, _I, _I]),
    'ynet_softargmax2d': (c_int, [_P, _I, _L, _I, _I, _P, _P, _L, _P]),
    'ynet_spatial_softmax': (c_int, [_P, _I, _L, _P, _P]),
    'ynet_expectation2d': (c_int, [_P, _I, _I, _I, _P, _P]),
    'ynet_sigmoid_select': (c_int, [_P, _I, _I, _L, POINTER(c_int32), _I, _F, _P, _P]),
    'ynet_sampling_prepare_workspace_bytes': (_L, [_I, _L]),
    'ynet_sampling_prepare': (c_int, [_P, _I, _L, _F, _P, _P, _P, _L, _P]),
    'ynet_multinomial_replacement': (c_int, [_P, _I, _L, _F, _P, _P, _P, _I, _P, _P, _P, _I, _P]),
    'ynet_multinomial_topk': (c_int, [_P, _P, _I, _L, _F, _P, _P, _I, _P, _P, _I, _P]),
    'ynet_counter_add': (c_int, [_P, c_uint64, _P]),
    'ynet_rng_uniform_f64': (c_int, [c_uint64, _P, c_uint64, _L, _P, _P]),
    'ynet_rng_exponential_f32': (c_int, [c_uint64, _P, c_uint64, _L, _P, _P]),
    'ynet_rng_choice': (c_int, [c_uint64, _P, c_uint64, _I, _I, _I, _P, _P]),
    'ynet_kmeans_batched': (c_int, [_P, _I, _I, _I, _P, _P, _I, _F, _I, _P, _P, _P, _P, _P]),
    'ynet_cws_waypoint_workspace_bytes': (_L, [_I, _I]),
    'ynet_cws_waypoint': (c_int, [_P, _I, _I, _I, _P, _I, _P, _F, _P, _F, _I, _P, _P, _L, _P]),
    'ynet_cws_waypoint_map': (c_int, [_P, _I, _I, _I, _P, _P, _F, _F, _F, _I, _P, _P]),
    'ynet_ade_fde': (c_int, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P]),
    'ynet_conv3x3_f32': (c_int, [POINTER(ConvSrc), _I, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    'ynet_conv1x1_f32': (c_int, [_P, _I, _I, _L, _P, _P, _I, _P, _P]),
    'ynet_predictor_softargmax_workspace_bytes': (_L, [_I, _I, _I, _I]),
    'ynet_predictor_softargmax_f32': (c_int, [_P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _L, _P]),
    'ynet_maxpool2x2_f32': (c_int, [_P, _L, _I, _I, _P, _P]),
    'ynet_upsample_bilinear2x_f32': (c_int, [_P, _L, _I, _I, _P, _P]),
    'ynet_lora_fold': (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    'ynet_tc_supported': (c_int, []),
    'ynet_tc_rasterize_im2col_c8': (c_int, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _I, _P]),
    'ynet_tc_rasterize_pyramid_c8': (c_int, [_P, _I, _I, _P, _I, _I, _I, _I, _I, POINTER(c_void_p), _I, _I, _I, _P, _P]),
    'ynet_tc_pack_f32_to_c8': (c_int, [_P, _I, _I, _I, _I, _L, _P, _I, _P]),
    'ynet_tc_unpack_c8_to_f32': (c_int, [_P, _I, _I, _I, _I, _I, _P, _P]),
    'ynet_tc_maxpool2x2': (c_int, [_P, _I, _I, _I, _I, _P, _P]),
    'ynet_tc_upsample2x': (c_int, [_P, _I, _I, _I, _I, _P, _P]),
    'ynet_tc_predictor_f32': (c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    'ynet_tc_packed_weight_bytes': (_L, [_I, _I, POINTER(c_int32), _I]),
    'ynet_tc_pack_weights': (c_int, [_P, _I, _I, POINTER(c_int32), POINTER(c_int32), _I, _P, _P]),
    'ynet_tc_conv1x1_f32': (c_int, [POINTER(TcSrc), _I, _I, _I, _I, _P, _P, _I, _P, _I, _P]),
    'ynet_tc_conv1x1_softargmax_workspace_bytes': (_L, [_I, _I, _I, _I]),
    'ynet_tc_conv1x1_softargmax': (c_int, [POINTER(TcSrc), _I, _I, _I, _I, _P, _P, _I, _P, _P, _L, _I, _P]),
    'ynet_tc_upconv_phase_weights': (c_int, [_P, _P, _I, _I, _P, _P, _P]),
    'ynet_tc_upconv_border_weight_bytes': (_L, [_I, _I, POINTER(c_int32)]),
    'ynet_tc_upconv_border_weights': (c_int, [_P, _I, _I, POINTER(c_int32), _P, _P]),
    'ynet_tc_upconv3x3': (c_int, [POINTER(TcSrc), POINTER(c_int32), _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _I, _P]),
    'ynet_tc_conv3x3': (c_int, [POINTER(TcSrc), _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _I, _P]),
    'ynet_tc_conv3x3_pred_softargmax': (c_int, [POINTER(TcSrc), _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P, _P, _L, _P]),
    'ynet_tc_pad_replicate': (c_int, [_P, _I, _I, _I, _I, _P, _P]),
    'ynet_tc_conv3x3_hilo': (c_int, [POINTER(TcSrc), _I, _I, _I, _I, _P, _I, _P, _I, _I, _I, _P]),
    'ynet_tc_conv3x3_split': (c_int, [POINTER(TcSrc), _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _I, _I, _I, _P]),
    'ynet_split_pack_f32': (c_int, [_P, _I, _I, _I, _I, _L, _P, _I, _I, _I, _P]),
    'ynet_split_pack_masked_f32': (c_int, [_P, _P, _I, _I, _I, _I, _P, _I, _P]),
    'ynet_split_unpack_f32': (c_int, [_P, _I, _I, _I, _I, _I, _P, _P]),
    'ynet_split_maxpool2x2': (c_int, [_P, _I, _I, _I, _I, _P, _P]),
    'ynet_split_upsample2x': (c_int, [_P, _I, _I, _I, _I, _P, _P]),
    'ynet_tc_rowconv_packed_weight_bytes': (_L, [_I]),
    'ynet_tc_rowconv_pack_weights': (c_int, [_P, _I, _I, _I, _P, _P]),
    'ynet_tc_rowconv3x3': (c_int, [POINTER(TcSrc), _I, POINTER(TcSrc), _I, _I, _I, _P, _P, _I, _I, _P, _I, _P]),
    'ynet_tc_wp_template_c8': (c_int, [_P, _I, _I, _P, _P, _P]),
    'ynet_tc_rowconv2_wp': (c_int, [POINTER(TcSrc), _I, POINTER(TcSrc), POINTER(TcWpSrc), _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _I,
                                    _P]),
    'ynet_tc_rowconv2_softargmax_workspace_bytes': (_L, [_I, _I, _I]),
    'ynet_tc_rowconv2_wp_pred_softargmax': (c_int, [POINTER(TcSrc), _I, POINTER(TcSrc), POINTER(TcWpSrc), _I, _I, _I, _P, _P, _P,
                                                    _P, _I, _P, _P, _I, _P, _P, _L, _P]),
    'ynet_tc_rowconv3x3_wp': (c_int, [POINTER(TcSrc), _I, POINTER(TcSrc), POINTER(TcWpSrc), _I, _I, _I, _P, _P, _I, _I, _P, _I,
                                      _P]),
    'ynet_tc_rowconv_softargmax_workspace_bytes': (_L, [_I, _I, _I]),
    'ynet_tc_rowconv3x3_pred_softargmax': (c_int, [POINTER(TcSrc), _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P, _P, _L, _P]),
    'ynet_scene_preprocess_u8': (c_int, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, POINTER(c_double),
                                         POINTER(c_double), _P, _P, _P]),
    'ynet_scene_preprocess_oriented_u8': (c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, POINTER(c_double),
                                                  POINTER(c_double), _P, _P, _P]),
    'ynet_scene_onehot_u8': (c_int, [_P, _I, _I, _I, _I, _I, _I, c_double, _I, _P, _P]),
    'ynet_bce_workspace_bytes': (_L, [_L]),
    'ynet_bce_logits_fwd_bwd': (c_int, [_P, _P, _L, _F, _P, _P, _P, _L, _P]),
    'ynet_conv3x3_dgrad_f32': (c_int, [_P, _P, _I, _I, _I, _P, _I, _I, _P, _P]),
    'ynet_conv3x3_wgrad_workspace_bytes': (_L, [_I, _I, _I, _I, _I]),
    'ynet_conv3x3_wgrad_f32': (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _L, _P]),
    'ynet_maxpool2x2_bwd_f32': (c_int, [_P, _P, _L, _I, _I, _P, _P]),
    'ynet_upsample_bilinear2x_bwd_f32': (c_int, [_P, _L, _I, _I, _P, _P]),
    'ynet_lora_grad': (c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    'ynet_adam_step': (c_int, [_P, _P, _P, _P, _L, _I, _F, _F, _F, _F, _F, _P]),
}


def lib_path():
    return _build.LIB_PATH


def load(build_if_missing=True):
    """Load (building first if sources are newer and nvcc exists) and type the C ABI."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if build_if_missing and _build.find_nvcc() is not None and _build.needs_build():
        _build.build(verbose=False)
    if not os.path.exists(path):
        raise YnetError(f'{path} is missing and could not be built (nvcc not found). '
                        'motion_style_transfer_b200 has no CPU fallback.')
    lib = ctypes.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)      # AttributeError = symbol missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().ynet_last_error_string()
        raise YnetError(f'{what} failed with code {rc}: {msg.decode() if msg else ""}')


def exported_names():
    return list(_PROTOS)
