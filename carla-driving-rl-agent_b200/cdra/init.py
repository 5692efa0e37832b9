"""Keras-default parameter initialisation for the flat arenas (what `CARLANetwork.__init__` gets from
building its Keras layers, core/networks.py:150-176): glorot-uniform kernels, glorot-uniform biases where
the reference passes `bias_initializer='glorot_uniform'` (every Dense and GRU; Conv2D/DepthwiseConv2D and
the alpha/beta heads keep Keras' zero bias), orthogonal GRU recurrent kernels (Keras `recurrent_initializer='orthogonal'`;
the reference only overrides the bias initialiser, core/networks.py:47-50), BatchNorm gamma 1 / beta 0 / moving mean 0 /
variance 1."""
import math

import torch


def _glorot(shape, fan_in, fan_out, gen, device):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, device=device) * 2.0 - 1.0) * lim


def _fans(name, shape):
    if name.endswith('dw.w'):                       # depthwise [3,3,C] == Keras [3,3,C,1]
        return 9 * shape[2], 9
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = 1
    for s in shape[:-2]:
        rf *= s
    return shape[-2] * rf, shape[-1] * rf


ZERO_BIAS_PREFIXES = ('tower.', 'alpha.', 'beta.')


def _orthogonal(shape, gen, device):
    """Keras Orthogonal initializer [lib] on the full [units, 3*units] recurrent matrix: QR of a normal matrix of the
    transposed (tall) shape, signs fixed by diag(R), transposed back."""
    rows, cols = shape
    a = torch.randn(max(rows, cols), min(rows, cols), generator=gen, device=device, dtype=torch.float32)
    q, r = torch.linalg.qr(a.double().cpu())
    q = q * torch.sign(torch.diagonal(r))
    if rows < cols:
        q = q.t()
    return q.to(torch.float32).to(device).contiguous()


def init_arena(arena, state, seed):
    gen = torch.Generator(device=arena.flat.device).manual_seed(seed)
    dev = arena.flat.device
    for name, shape in zip(arena.names, arena.shapes):
        v = arena.view(name)
        if name.endswith('.g'):
            v.fill_(1.0)
        elif name.endswith('.be'):
            v.zero_()
        elif name.endswith('.b') and name.startswith(ZERO_BIAS_PREFIXES):
            v.zero_()
        elif name.startswith('gru.') and name.endswith('.r'):
            v.copy_(_orthogonal(shape, gen, dev))
        else:
            fi, fo = _fans(name, shape)
            v.copy_(_glorot(shape, fi, fo, gen, dev))
    for name in state.names:
        state.view(name).fill_(1.0 if name.endswith('.mv') else 0.0)


def init_engine(eng, seed=42):
    """Same seed -> same weights on every data-parallel rank."""
    init_arena(eng.dyn, eng.dyn_state, seed)
    init_arena(eng.pol, eng.pol_state, seed + 1)
    init_arena(eng.val, eng.val_state, seed + 2)
