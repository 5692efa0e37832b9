"""The C-ABI shared library loads without a GPU and exports every symbol include/cdra.h declares."""
import ctypes
import os
import re

import pytest

from tests.conftest import PKG, ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'cdra.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(cdra_[a-z0-9_]+)\s*\(', src)))


def test_header_and_binding_agree(built_libs):
    from cdra import _lib
    assert _declared_symbols() == _lib.exported_symbols()


def test_cuda_library_exports_every_symbol(built_libs):
    path = os.path.join(PKG, 'cdra', 'libcdra.so')
    if not os.path.exists(path):
        pytest.skip('nvcc not available on this box')
    lib = ctypes.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(lib, s), s
    from cdra import _lib
    l2 = _lib.load()
    assert l2.cdra_version() >= 100
    # host-only calls work without a device
    cfg = _lib.Config(4, 90, 120, _lib.BF16, 1)
    plan = ctypes.c_void_p()
    assert l2.cdra_plan_create(ctypes.byref(cfg), ctypes.byref(plan)) == 0
    assert l2.cdra_arena_size(plan, _lib.ARENA_DYN_PARAMS) == 2128450
    assert l2.cdra_arena_size(plan, _lib.ARENA_DYN_STATE) == 16564
    assert l2.cdra_arena_size(plan, _lib.ARENA_POL_PARAMS) == 270470
    assert l2.cdra_arena_size(plan, _lib.ARENA_VAL_PARAMS) == 269828
    assert l2.cdra_plan_workspace_bytes(plan) > 0
    # every BASELINE geometry plans on the host: config 2 (90x120), config 4 (180x240: stage-1 frames exceed shared memory,
    # the depthwise kernels band them), odd sizes
    sizes = {}
    for (h, w) in ((90, 120), (180, 240), (91, 123)):
        pl = ctypes.c_void_p()
        assert l2.cdra_plan_create(ctypes.byref(_lib.Config(2, h, w, _lib.BF16, 1)), ctypes.byref(pl)) == 0
        sizes[(h, w)] = l2.cdra_plan_workspace_bytes(pl)
        off, dims, es = ctypes.c_int64(), (ctypes.c_int32 * 4)(), ctypes.c_int32()
        assert l2.cdra_plan_tensor(pl, b'tower.stem', ctypes.byref(off), dims, ctypes.byref(es)) == 0
        assert list(dims) == [8, (h - 3) // 2 + 1, (w - 3) // 2 + 1, 24] and es.value == 2
        l2.cdra_plan_destroy(pl)
    assert sizes[(180, 240)] > 1.5 * sizes[(90, 120)]       # 4x the activations; the fixed GEMM-operand buffers do not scale
    # debug switches: known keys only
    assert l2.cdra_debug_set(b'tc', -1) == 0 and l2.cdra_debug_set(b'fwd_tc', -1) == 0 and l2.cdra_debug_set(b'nope', 1) == -1
    # shape validation of the self tests happens before any device work
    assert l2.cdra_debug_umma_selftest(None, None, None, 64, 128, 64, None) == -1
    assert l2.cdra_debug_umma_selftest_k(ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), 100, 64, 64, None) == -1
    bad = _lib.Config(0, 90, 120, 0, 1)
    assert l2.cdra_plan_create(ctypes.byref(bad), ctypes.byref(ctypes.c_void_p())) == -2
    assert b'batch' in l2.cdra_last_error()
    l2.cdra_plan_destroy(plan)


def test_product_loader_has_no_cpu_fallback(built_libs):
    from cdra import _lib
    from cdra.engine import Engine
    with pytest.raises(_lib.CdraError):
        Engine(2, 42, 58, device='cpu')          # the CUDA library refuses CPU tensors; only tests may ask for the emulator
