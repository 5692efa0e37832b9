"""Layer-by-layer restatement of the tower with bf16 STORAGE (oracle; test infrastructure).

The perf mode of the library stores every raw conv output and every activation gradient in
bf16, feeds bf16 operands to the tensor cores and accumulates in fp32.  A 50-layer training-mode BatchNorm tower
amplifies rounding noise, so an end-to-end comparison with the fp64 oracle cannot be tight.  These functions restate
ONE layer at a time with the rounding applied at exactly the library's storage points:

    activated input  a = bf16(fma(raw, scale, shift)) (then ReLU6)     -- the consumer applies the producer's BatchNorm
    conv operands    bf16(W) for the pointwise convs, fp32 W for the depthwise convs
    stored output    bf16(fp32 accumulate + bias)
    backward         dR = bf16(scale * (dz - S1/n - xhat * S2/n)), dz = dA * [0 < z < 6]; stored gradients in bf16

so that a test can feed the library's own stored inputs of a layer (read back through cdra_debug_export) and hold
the layer's outputs / gradients to bf16 resolution.  The math follows core/architectures.py:44-57,120-173 (Conv2D /
DepthwiseConv2D / BatchNormalization / ReLU(max 6) / channel shuffle) and tape.gradient of them
(core/carla_agent.py:361-365); frames are ordered [slice t][sample b] like the library (one BatchNorm call per slice).
"""
import torch
import torch.nn.functional as F

from . import model, spec

T = spec.TIME_HORIZON


def bf16r(x):
    """round-to-nearest-even to bf16, result in the input dtype"""
    return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)


def f32r(x):
    return x.to(torch.float32).to(x.dtype)


def per_slice(x):
    """[4B, ...] -> [4, B, ...]"""
    return x.reshape(T, x.shape[0] // T, *x.shape[1:])


def bn_tables(raw, gamma, beta):
    """(scale, shift, mean, inv_std), each [4][C], from the stored raw values of one tensor (v2_common.cuh bn_from_sums):
    biased batch variance per time slice, eps 1e-3 (Keras BatchNormalization defaults [lib])."""
    x = per_slice(raw.double())
    dims = tuple(range(1, x.dim() - 1))
    mean = x.mean(dim=dims)
    var = (x * x).mean(dim=dims) - mean * mean
    inv = f32r(torch.rsqrt(var.clamp_min(0.0) + spec.BN_EPS))
    scale = f32r(gamma.double() * inv)
    shift = f32r(beta.double() - f32r(mean) * scale)
    return scale, shift, f32r(mean), inv


def _bc(tab, x):
    """[4][C] table -> broadcastable against [4, B, ..., C]"""
    return tab.reshape(T, *([1] * (x.dim() - 2)), tab.shape[-1])


def activate(raw, scale, shift, clamp):
    """what a consumer kernel stages: bf16(fma(raw, scale, shift)), then ReLU6 on the rounded value (affine8)"""
    x = per_slice(raw.double())
    z = bf16r(f32r(x * _bc(scale, x) + _bc(shift, x)))
    if clamp:
        z = z.clamp(0.0, 6.0)
    return z.reshape(raw.shape)


def pw_forward(act_in, w, b):
    """Conv2D 1x1 (core/architectures.py:130,134,140,170): bf16 operands, wide accumulate, bf16 store"""
    return bf16r(act_in.double() @ bf16r(w.double()) + b.double())


def dw_forward(act_in, w, b, stride):
    """DepthwiseConv2D 3x3 SAME (:132,138): bf16 activations, fp32 weights, bf16 store"""
    return bf16r(model.depthwise3x3(act_in.double(), w.double(), b.double(), stride))


def bn_backward(dA, raw, scale, shift, mean, inv, clamp, S1=None, S2=None):
    """BatchNorm(+ReLU6) backward of a stored tensor (v2_bwd.cuh bnbwd_apply): returns (dR as the kernels stage it (bf16),
    S1 [4][C], S2 [4][C]).  When S1 / S2 are given (the library's own sums) they are used instead of the recomputed ones."""
    x = per_slice(raw.double()); d = per_slice(dA.double())
    n = x.numel() // (T * x.shape[-1])
    z = f32r(x * _bc(scale, x) + _bc(shift, x))
    dz = torch.where((z > 0.0) & (z < 6.0), d, torch.zeros_like(d)) if clamp else d
    xhat = (x - _bc(mean, x)) * _bc(inv, x)
    dims = tuple(range(1, x.dim() - 1))
    s1 = dz.sum(dim=dims); s2 = (dz * xhat).sum(dim=dims)
    u1 = s1 if S1 is None else S1.double(); u2 = s2 if S2 is None else S2.double()
    dR = _bc(scale, x) * (dz - _bc(u1, x) / n - xhat * _bc(u2, x) / n)
    return bf16r(dR).reshape(raw.shape), s1, s2


def pw_backward(dR, act_in, w):
    """data gradient (bf16 store) and weight gradient (fp32) of a 1x1 conv from bf16 operands"""
    wb = bf16r(w.double())
    d_in = dR.double() @ wb.t()
    dW = act_in.double().reshape(-1, act_in.shape[-1]).t() @ dR.double().reshape(-1, dR.shape[-1])
    return d_in, dW


def dw_backward(dR, act_in, w, stride):
    """data gradient and weight gradient of the depthwise conv (autograd of the fp64 restatement)"""
    a = act_in.double().clone().requires_grad_(True)
    ww = w.double().clone().requires_grad_(True)
    y = model.depthwise3x3(a, ww, torch.zeros(w.shape[-1], dtype=torch.float64), stride)
    ga, gw = torch.autograd.grad(y, [a, ww], dR.double())
    return ga, gw


def unshuffle(x):
    """inverse of channel_shuffle: concat[2i + g] = out[g * C/2 + i]"""
    c = x.shape[-1]
    lead = x.shape[:-1]
    return x.reshape(*lead, 2, c // 2).transpose(-1, -2).reshape(*lead, c)


def maxpool_first(act):
    """MaxPooling2D(3, 2, 'same') with the FIRST maximum in scan order as the winner (v2_stem.cuh pool_fwd_kernel; TF's
    CPU max-pool gradient picks the same element [lib]).  Returns (pooled NHWC, flat winner index into H*W per channel)."""
    xc = model._pad_same_nchw(act.permute(0, 3, 1, 2), 3, 2, value=float('-inf'))
    y, idx = F.max_pool2d(xc, 3, 2, return_indices=True)
    return y.permute(0, 2, 3, 1), idx, xc.shape[-2:]


def maxpool_backward(dpool, idx, padded_hw, act_shape):
    """scatter d pool through the winner positions -> d act [F, H, W, C]"""
    Fn, H, W, Cn = act_shape
    ph, pw = padded_hw
    g = torch.zeros(Fn, Cn, ph * pw, dtype=torch.float64)
    g.scatter_add_(2, idx.reshape(Fn, Cn, -1), dpool.double().permute(0, 3, 1, 2).reshape(Fn, Cn, -1))
    g = g.reshape(Fn, Cn, ph, pw)
    pt, _ = spec.same_pad(H, 3, 2); pl, _ = spec.same_pad(W, 3, 2)
    return g[:, :, pt:pt + H, pl:pl + W].permute(0, 2, 3, 1)
