"""Map the reference's shipped TF checkpoints (weights/stage-*/) onto the oracle's parameter names.

Keras stores a functional model's variables as `layer_with_weights-N` in topological order; layers of
equal depth are ordered shortcut-before-branch.  Shapes are asserted at every step,
so a wrong ordering cannot load silently.  Test infrastructure / checkpoint-import row.
"""
import numpy as np

from . import spec, tf_bundle


def _conv(layer):
    k = layer.get('kernel', layer.get('depthwise_kernel'))
    return k, layer['bias']


def dynamics_layer_order():
    """[(oracle prefix, kind)] in checkpoint layer order; kind in conv|dw|bn|dense|gru."""
    order = [('tower.stem', 'conv'), ('tower.stem', 'bn')]
    for name, stride, cin, c in spec.tower_units():
        if stride == 2:
            order += [(name + '.pw1', 'conv'), (name + '.pw1', 'bn'), (name + '.scdw', 'dw'), (name + '.dw', 'dw'),
                      (name + '.scdw', 'bn'), (name + '.dw', 'bn'), (name + '.scpw', 'conv'), (name + '.pw2', 'conv'),
                      (name + '.scpw', 'bn'), (name + '.pw2', 'bn')]
        else:
            order += [(name + '.pw1', 'conv'), (name + '.pw1', 'bn'), (name + '.dw', 'dw'), (name + '.dw', 'bn'),
                      (name + '.pw2', 'conv'), (name + '.pw2', 'bn')]
    feats = [f for f, _ in spec.FEATURES]
    order += [('tower.head', 'conv')] + [(f'feat.{f}.d1', 'dense') for f in feats] + [('tower.head', 'bn')]
    order += [(f'feat.{f}.d1', 'bn') for f in feats] + [(f'feat.{f}.d2', 'dense') for f in feats]
    order += [(f'feat.{f}.d2', 'bn') for f in feats]
    order += [(f'gru.{g}', 'gru') for g, _, _ in spec.GRU_UNITS]
    order += [('trunk.bn', 'bn'), ('trunk.dense', 'dense')]
    return order


def head_layer_order(kind):
    heads = ('alpha', 'beta', 'similarity', 'speed') if kind == 'policy' else ('base', 'exp', 'speed', 'similarity')
    return [('bn1', 'bn'), ('d1', 'dense'), ('bn2', 'bn'), ('d2', 'dense')] + [(h, 'dense') for h in heads]


def _assign(layers, order, shapes):
    out = {}
    assert len(layers) == len(order), (len(layers), len(order))
    for i, (prefix, kind) in enumerate(order):
        lay = layers[i]
        if kind == 'bn':
            out[prefix + '.g'], out[prefix + '.be'] = lay['gamma'], lay['beta']
            out[prefix + '.mm'], out[prefix + '.mv'] = lay['moving_mean'], lay['moving_variance']
        elif kind == 'gru':
            out[prefix + '.k'], out[prefix + '.r'], out[prefix + '.b'] = lay['cell/kernel'], lay['cell/recurrent_kernel'], lay['cell/bias']
        else:
            k, b = _conv(lay)
            if kind == 'conv':
                k = k.reshape(k.shape[-2], k.shape[-1]) if k.shape[0] == 1 else k     # [1,1,K,N] -> [K,N]
            elif kind == 'dw':
                k = k.reshape(3, 3, k.shape[2])                                        # [3,3,C,1] -> [3,3,C]
            out[prefix + '.w'], out[prefix + '.b'] = k, b
    for name, shape in shapes.items():
        assert tuple(out[name].shape) == tuple(shape), (name, out[name].shape, shape)
    return out


def load_reference_checkpoint(folder):
    """-> (dynamics, policy, value) dicts of numpy arrays keyed by oracle/spec.py names."""
    res = []
    for fname, order, pspec in (('dynamics_model', dynamics_layer_order(), spec.dynamics_params()),
                                ('policy_net', head_layer_order('policy'), spec.head_params('policy')),
                                ('value_net', head_layer_order('value'), spec.head_params('value'))):
        layers = tf_bundle.layer_variables(f'{folder}/{fname}')
        layers = [layers[i] for i in sorted(layers)]
        shapes = {n: s for n, s, _ in pspec}
        res.append(_assign(layers, order, shapes))
    return tuple(res)


def index_summary(folder):
    """{model: [[variable key, shape], ...]} straight from the .index files (golden architecture pin)."""
    out = {}
    for fname in ('dynamics_model', 'policy_net', 'value_net'):
        idx = tf_bundle.read_index(f'{folder}/{fname}')
        out[fname] = sorted([k, list(v['shape'])] for k, v in idx.items() if k.endswith('VARIABLE_VALUE') and k.startswith('layer_with_weights'))
    return out
