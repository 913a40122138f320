"""ctypes binding of libnsw_b200.so (include/nsw.h).

The product path has no CPU fallback: if the shared library cannot be loaded the
import of any engine raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

NSW_MAX_FLOWS = 8
NSW_MAX_DECONV = 4

LOSS = {'logistic': 0, 'gauss': 1, 'mol': 2, 'ce': 3}
ACT = {'tanh': 0, 'relu': 1, 'leaky_relu': 2}
ENGINE = {'ffma': 0, 'tc': 1, 'tc2': 2, 'tc3': 3}

ERRORS = {-1: 'NSW_EINVAL', -2: 'NSW_ECUDA', -3: 'NSW_EMISSING', -4: 'NSW_ETIMEOUT', -5: 'NSW_ERANGE'}


class NswError(RuntimeError):
    pass


class nsw_tensor(C.Structure):
    _fields_ = [('name', C.c_char_p), ('data', C.POINTER(C.c_float)), ('ndim', C.c_int32),
                ('shape', C.c_int64 * 4)]


class nsw_iaf_config(C.Structure):
    _fields_ = [('num_flows', C.c_int32), ('num_iaf_layers', C.c_int32 * NSW_MAX_FLOWS),
                ('num_stages', C.c_int32), ('filter_length', C.c_int32), ('width', C.c_int32),
                ('deconv_width', C.c_int32), ('num_mel', C.c_int32), ('num_deconv', C.c_int32),
                ('deconv_filter', C.c_int32 * NSW_MAX_DECONV),
                ('deconv_stride', C.c_int32 * NSW_MAX_DECONV), ('share_deconv', C.c_int32),
                ('loss_type', C.c_int32), ('upsample_act', C.c_int32), ('use_mu_law', C.c_int32),
                ('engine', C.c_int32)]


class nsw_wavenet_config(C.Structure):
    _fields_ = [('num_layers', C.c_int32), ('num_stages', C.c_int32),
                ('filter_length', C.c_int32), ('width', C.c_int32), ('gate_width', C.c_int32),
                ('skip_width', C.c_int32), ('out_width', C.c_int32), ('deconv_width', C.c_int32),
                ('num_mel', C.c_int32), ('num_deconv', C.c_int32),
                ('deconv_filter', C.c_int32 * NSW_MAX_DECONV),
                ('deconv_stride', C.c_int32 * NSW_MAX_DECONV), ('loss_type', C.c_int32),
                ('upsample_act', C.c_int32), ('use_mu_law', C.c_int32), ('engine', C.c_int32)]


_FP = C.POINTER(C.c_float)
_VP = C.c_void_p

# name -> (restype, argtypes); every symbol include/nsw.h declares
PROTOTYPES = {
    'nsw_version': (C.c_int, []),
    'nsw_last_error': (C.c_char_p, []),
    'nsw_kernel_launch_count': (C.c_uint64, []),
    'nsw_range_status': (C.c_int, [C.c_int32]),
    'nsw_crc32c': (C.c_uint32, [_VP, C.c_size_t, C.c_uint32]),
    'nsw_iaf_create': (C.c_int, [C.POINTER(nsw_iaf_config), C.POINTER(nsw_tensor), C.c_int32,
                                 C.c_int32, C.POINTER(_VP)]),
    'nsw_iaf_destroy': (None, [_VP]),
    'nsw_iaf_length': (C.c_int64, [_VP, C.c_int32]),
    'nsw_iaf_forward_device': (C.c_int, [_VP, _VP, _VP, C.c_uint64, C.c_int32, C.c_int32,
                                         C.c_int32, _VP, _VP, _VP, _VP, _VP, _VP]),
    'nsw_iaf_forward_host': (C.c_int, [_VP, _VP, _VP, C.c_uint64, C.c_int32, C.c_int32,
                                       C.c_int32, _VP, _VP, _VP, _VP, _VP]),
    'nsw_iaf_deconv_device': (C.c_int, [_VP, C.c_int32, _VP, C.c_int32, C.c_int32, _VP, _VP]),
    'nsw_iaf_set_tap': (C.c_int, [_VP, C.c_int32, C.c_int32, _VP]),
    'nsw_iaf_workspace_bytes': (C.c_size_t, [_VP]),
    'nsw_iaf_set_profiling': (C.c_int, [_VP, C.c_int32]),
    'nsw_iaf_last_timing': (C.c_int, [_VP, C.POINTER(C.c_float * 5)]),
    'nsw_fastgen_create': (C.c_int, [C.POINTER(nsw_wavenet_config), C.POINTER(nsw_tensor),
                                     C.c_int32, C.c_int32, C.POINTER(_VP)]),
    'nsw_fastgen_destroy': (None, [_VP]),
    'nsw_fastgen_encode_device': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP]),
    'nsw_fastgen_encode_host': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP]),
    'nsw_fastgen_run_device': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP, C.c_uint64, _VP,
                                         _VP, _VP]),
    'nsw_fastgen_run_host': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP, C.c_uint64, _VP,
                                       _VP]),
    'nsw_fastgen_last_timing': (C.c_int, [_VP, C.POINTER(C.c_float)]),
    'nsw_fastgen_cond_vars_device': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP]),
    'nsw_fastgen_cond_vars_host': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP]),
    'nsw_fastgen_set_noise': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    'nsw_fastgen_gn_pack_host': (C.c_int, [C.POINTER(nsw_wavenet_config), C.POINTER(nsw_tensor),
                                           C.c_int32, _VP, C.c_int64, _VP, _VP,
                                           C.POINTER(C.c_int64)]),
    'nsw_teacher_create': (C.c_int, [C.POINTER(nsw_wavenet_config), C.POINTER(nsw_tensor), C.c_int32,
                                     C.c_int32, C.POINTER(_VP)]),
    'nsw_teacher_destroy': (None, [_VP]),
    'nsw_teacher_forward_device': (C.c_int, [_VP, _VP, _VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP]),
    'nsw_teacher_forward_host': (C.c_int, [_VP, _VP, _VP, C.c_int32, C.c_int32, C.c_int32, _VP]),
    'nsw_teacher_last_timing': (C.c_int, [_VP, C.POINTER(C.c_float)]),
    'nsw_mol_score_device': (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_uint64, C.c_int32, C.c_int32,
                                       C.c_int32, C.POINTER(C.c_double * 3), _VP]),
    'nsw_gauss_kl_device': (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_int32, C.c_int32, C.POINTER(C.c_double * 3),
                                      _VP]),
    'nsw_mel_create': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, C.c_float,
                                 C.c_float, C.POINTER(_VP)]),
    'nsw_mel_destroy': (None, [_VP]),
    'nsw_mel_frames': (C.c_int, [_VP, C.c_int32]),
    'nsw_mel_device': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP]),
    'nsw_mel_host': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP]),
    'nsw_mel_set_framing': (C.c_int, [_VP, C.c_int32, C.c_int32]),
    'nsw_stft_mag_device': (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP]),
    'nsw_power_loss_device': (C.c_int, [_VP, _VP, C.c_int32, _VP, C.c_int32, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_double * 3), _VP]),
    'nsw_fastgen_pack_host': (C.c_int, [C.POINTER(nsw_wavenet_config), C.POINTER(nsw_tensor),
                                        C.c_int32, _VP, C.c_int64, _VP, _VP,
                                        C.POINTER(C.c_int64)]),
    'nsw_flow_pair_plan_host': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                          C.POINTER(C.c_int64)]),
    'nsw_conv_gemm_device': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'nsw_flow_plan_host': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     _VP, C.c_int64, C.POINTER(C.c_int64)]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load(build_if_stale=True):
    """Load (building first when sources are newer and nvcc exists) the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get('NSW_LIB')      # timing experiments: an alternative build of the same sources
    if override:
        _build.LIB_PATH = override
    elif build_if_stale and _build.is_stale():
        _build.build_library()
    if not os.path.exists(_build.LIB_PATH):
        raise NswError('libnsw_b200.so is missing and could not be built; '
                       'there is no CPU fallback for the generation path')
    lib = C.CDLL(_build.LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().nsw_last_error().decode('utf-8', 'replace')
        raise NswError('{} ({}): {}'.format(ERRORS.get(rc, 'error'), rc, msg))


def make_tensors(weights):
    """dict name -> np.ndarray  =>  (ctypes array of nsw_tensor, keepalive list)."""
    keep = []
    arr = (nsw_tensor * len(weights))()
    for i, (name, v) in enumerate(weights.items()):
        a = np.ascontiguousarray(v, dtype=np.float32)
        b = name.encode('utf-8')
        keep.append((a, b))
        arr[i].name = b
        arr[i].data = a.ctypes.data_as(_FP)
        arr[i].ndim = min(a.ndim, 4)
        shape = list(a.shape) if a.ndim <= 4 else [int(np.prod(a.shape[:-3]))] + list(a.shape[-3:])
        for d in range(4):
            arr[i].shape[d] = shape[d] if d < len(shape) else 1
    return arr, keep


def ptr(x):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()
