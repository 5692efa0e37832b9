"""ctypes binding of libcdra (include/cdra.h).  PyTorch is used only for device memory / streams.

The product path loads `cdra/libcdra.so` (nvcc, sm_100a) and fails loudly when it is missing — there
is no CPU fallback and no switch that selects one.  (The CPU test-suite compiles the same sources against a SIMT
emulator and binds that library itself, under `tests/emu/`, through `bind()`.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libcdra.so')

F32, BF16 = 0, 1
ARENA_DYN_PARAMS, ARENA_DYN_STATE, ARENA_POL_PARAMS, ARENA_POL_STATE, ARENA_VAL_PARAMS, ARENA_VAL_STATE = range(6)


class CdraError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [('batch', C.c_int32), ('height', C.c_int32), ('width', C.c_int32), ('dtype', C.c_int32),
                ('image_u8', C.c_int32)]


_P = C.c_void_p
_SIGNATURES = {
    'cdra_last_error': (C.c_char_p, []),
    'cdra_version': (C.c_int, []),
    'cdra_plan_create': (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    'cdra_plan_destroy': (None, [_P]),
    'cdra_plan_workspace_bytes': (C.c_size_t, [_P]),
    'cdra_arena_size': (C.c_int64, [_P, C.c_int]),
    'cdra_arena_num_tensors': (C.c_int, [_P, C.c_int]),
    'cdra_arena_tensor': (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'cdra_plan_tensor': (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'cdra_debug_export': (C.c_int, [_P, C.c_char_p, _P, _P, C.POINTER(C.c_int32), _P]),
    'cdra_debug_gemm': (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    'cdra_debug_umma_selftest_k': (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    'cdra_debug_umma_selftest': (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    'cdra_debug_set': (C.c_int, [C.c_char_p, C.c_int]),
    'cdra_debug_timeline': (C.c_int, [_P]),
    'cdra_debug_stem_backward': (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P]),
    'cdra_dynamics_forward': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P]),
    'cdra_dynamics_backward': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'cdra_policy_head_loss_fwd_bwd': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_float, C.c_float, C.c_int,
                                                C.c_float, _P, _P, _P, _P, _P, _P]),
    'cdra_value_head_loss_fwd_bwd': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_float, _P, _P, _P, _P, _P, _P]),
    'cdra_gae': (C.c_int, [_P, _P, _P, C.c_double, C.c_double, C.c_float, C.c_int, C.c_int, _P, _P, _P]),
    'cdra_clip_adam': (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_int64, C.c_float, _P, _P]),
    'cdra_grad_norms': (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_float, _P, _P]),
    'cdra_gather_rows': (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P]),
    'cdra_gather_rows_multi': (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int64, _P]),
    'cdra_comm_unique_id': (C.c_int, [_P]),
    'cdra_comm_create': (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    'cdra_comm_destroy': (None, [_P]),
    'cdra_allreduce_grads': (C.c_int, [_P, _P, C.c_int64, _P]),
    'cdra_augment': (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    'cdra_launch_count': (C.c_int64, []),
    'cdra_profile_enable': (None, [C.c_int]),
    'cdra_profile_reset': (None, []),
    'cdra_profile_report': (C.c_int, [C.c_char_p, C.c_int]),
}

_loaded = {}


def bind(path):
    """dlopen `path` and declare every symbol of include/cdra.h on it (AttributeError if one is not exported)."""
    if path in _loaded:
        return _loaded[path]
    if not os.path.exists(path):
        raise CdraError(f'{path} is missing: build it with `python carla-driving-rl-agent_b200/build.py` '
                        '(there is no fallback path)')
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _loaded[path] = lib
    return lib


def load():
    """The sm_100a library of the product path."""
    return bind(LIB_PATH)


def exported_symbols():
    return sorted(_SIGNATURES)


def check(lib, code, what=''):
    if code != 0:
        raise CdraError(f'{what} failed ({code}): {lib.cdra_last_error().decode()}')


def ptr(t):
    """Raw pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
