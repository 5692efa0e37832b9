"""Host side of the on-device image augmentation (include/cdra.h `cdra_augment`): the parameter struct, the draw of
the per-call random scalars with the reference's chance gates (core/carla_agent.py:527-579), and the launch wrapper.

The reference draws with TF's stateful RNG; here the per-call scalars come from a `numpy.random.Generator` and the
per-pixel draws from the library's counter-based hash (seeded per call), so a call is reproducible from (generator
state, seed) and the oracle can be fed the identical numbers.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

EPS = float(np.finfo(np.float32).eps)          # utils.EPSILON (rl/utils.py:24-25)


class AugmentParams(C.Structure):
    _fields_ = [('seed', C.c_uint32), ('jitter', C.c_int32),
                ('brightness', C.c_float), ('contrast', C.c_float), ('saturation', C.c_float), ('hue', C.c_float),
                ('blur_size', C.c_int32), ('blur_kernel', C.c_float * 75),
                ('salt_pepper', C.c_int32), ('sp_amount', C.c_float),
                ('gauss_noise', C.c_int32), ('gn_amount', C.c_float), ('gn_std', C.c_float),
                ('normalize', C.c_int32), ('group', C.c_int32), ('eps', C.c_float),
                ('cutout_size', C.c_int32), ('cutout_cell', C.c_int32), ('dropout_size', C.c_int32)]


def identity_params(group=1):
    p = AugmentParams()
    p.contrast = 1.0; p.saturation = 1.0; p.group = group; p.eps = EPS
    return p


def draw_params(rng: np.random.Generator, alpha: float, group: int = 1):
    """One call of `augment_fn` with intensity alpha > 0: every `tf_chance() < p * alpha` gate and every scalar the gated
    op draws.  Returns (AugmentParams, dropout_mask uint8 [81, 81] or None)."""
    p = identity_params(group)
    p.seed = int(rng.integers(0, 2 ** 32, dtype=np.uint64))
    mask = None
    if rng.random() < alpha:                               # color_jitter(strength=alpha)   simclr.py:44-50
        p.jitter = 1
        p.brightness = float(rng.uniform(-0.2 * alpha, 0.2 * alpha))
        p.contrast = float(rng.uniform(1.0 - 0.8 * alpha, 1.0 + 0.8 * alpha))
        p.saturation = float(rng.uniform(1.0 - 0.8 * alpha, 1.0 + 0.8 * alpha))
        p.hue = float(rng.uniform(-0.2 * alpha, 0.2 * alpha))
    if rng.random() < 0.25 * alpha:                        # tf_gaussian_blur: kernel ~ N(1, 0.25), size 3 | 5
        size = 3 if rng.random() >= 0.5 else 5
        p.blur_size = size
        k = rng.normal(1.0, 0.25, size=(size, size, 3)).astype(np.float32).ravel()
        for i, v in enumerate(k):
            p.blur_kernel[i] = float(v)
    if rng.random() < 0.2 * alpha:                         # tf_salt_and_pepper_batch(amount=0.1)
        p.salt_pepper = 1; p.sp_amount = 0.1
    if rng.random() < 0.33 * alpha:                        # tf_gaussian_noise_batch(amount=0.10, std=0.075)
        p.gauss_noise = 1; p.gn_amount = 0.10; p.gn_std = 0.075
    p.normalize = 1                                        # tf_normalize_batch (always, when alpha > 0)
    if rng.random() < 0.15 * alpha:                        # tf_cutout_batch(size=6): the arg-max cell of a 6 x 6 normal draw
        p.cutout_size = 6
        p.cutout_cell = int(np.argmax(rng.normal(size=36)))
    if rng.random() < 0.15 * alpha:                        # tf_coarse_dropout_batch(size=81, amount=0.04)
        p.dropout_size = 81
        mask = (rng.random((81, 81)) < 1.0 - 0.04).astype(np.uint8)
    return p, mask


def augment(images: torch.Tensor, params: AugmentParams, dropout_mask=None, lib=None, out=None):
    """images: CUDA tensor [..., H, W, 3], uint8 (value / 255 is augmented) or float32 -> float32 tensor of the same shape."""
    lib = lib or _lib.load()
    assert images.is_cuda and images.is_contiguous() and images.shape[-1] == 3 and images.dtype in (torch.uint8, torch.float32)
    H, W = int(images.shape[-3]), int(images.shape[-2])
    frames = images.numel() // (H * W * 3)
    out = torch.empty(images.shape, dtype=torch.float32, device=images.device) if out is None else out
    groups = (frames + max(1, params.group) - 1) // max(1, params.group)
    scratch = torch.empty(3 * frames + 2 * groups + 4, dtype=torch.int32, device=images.device)
    dm = None
    if params.dropout_size > 0:
        dm = torch.as_tensor(np.ascontiguousarray(dropout_mask, dtype=np.uint8)).to(images.device)
        assert dm.numel() == params.dropout_size ** 2
    stream = C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)
    _lib.check(lib, lib.cdra_augment(_lib.ptr(images), 1 if images.dtype == torch.uint8 else 0, frames, H, W, C.byref(params),
                                     _lib.ptr(dm), _lib.ptr(out), _lib.ptr(scratch), stream), 'augment')
    return out
