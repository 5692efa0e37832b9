"""Checkpoint I/O of the three networks (reference: core/networks.py:297-310, rl/agents/agents.py:195-203).

Two on-disk formats are understood by `load_model`:

* the reference's own checkpoints -- TensorFlow "tensor bundle" files written by `keras.Model.save_weights`
  (`weights/stage-*/{dynamics_model,policy_net,value_net}.index` + `.data-0000N-of-0000M`), so the agents the
  reference ships load straight into the arenas (`CARLAgent(load=True, load_full=...)`);
* this library's `<prefix>.npz` (one array per arena tensor), written by `save_model`.

Bundle format: `.index` is an uncompressed LevelDB table (prefix-compressed keys in restart blocks, one index block,
48-byte footer ending in the table magic); each value is a BundleEntryProto {1 dtype, 2 shape, 3 shard, 4 offset,
5 size}; tensors are raw little-endian arrays inside the data shards.  Keras names a functional model's variables
`layer_with_weights-<i>/<attr>/.ATTRIBUTES/VARIABLE_VALUE` with i in topological order, which differs from the
construction order of the arenas inside every ShuffleNet unit: `keras_layer_sequence` restates that order and every
assignment is checked by shape, so a wrong order cannot load silently.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
NP_DTYPE = {1: '<f4', 2: '<f8', 3: '<i4', 9: '<i8'}
SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'


class _Cursor:
    __slots__ = ('b', 'i')

    def __init__(self, b, i=0):
        self.b, self.i = b, i

    def varint(self):
        v = s = 0
        while True:
            c = self.b[self.i]
            self.i += 1
            v |= (c & 0x7F) << s
            if c < 0x80:
                return v
            s += 7

    def take(self, n):
        out = self.b[self.i:self.i + n]
        self.i += n
        return out

    def done(self):
        return self.i >= len(self.b)


def _table_block(buf, off, size):
    """(key, value) pairs of one LevelDB block: entries are (shared, unshared, value_len, key tail, value)."""
    blk = memoryview(buf)[off:off + size]
    nrestart = struct.unpack_from('<I', blk, len(blk) - 4)[0]
    cur, stop, key = _Cursor(blk), len(blk) - 4 * (nrestart + 1), b''
    while cur.i < stop:
        shared, unshared, vlen = cur.varint(), cur.varint(), cur.varint()
        key = key[:shared] + bytes(cur.take(unshared))
        yield key, bytes(cur.take(vlen))


def _proto_fields(msg):
    """protobuf wire decoding, enough for BundleEntryProto / TensorShapeProto: {field: [values]}"""
    cur, out = _Cursor(msg), {}
    while not cur.done():
        tag = cur.varint()
        kind = tag & 7
        if kind == 0:
            v = cur.varint()
        elif kind == 2:
            v = bytes(cur.take(cur.varint()))
        elif kind == 5:
            v = struct.unpack('<I', cur.take(4))[0]
        elif kind == 1:
            v = struct.unpack('<Q', cur.take(8))[0]
        else:
            raise ValueError(f'protobuf wire type {kind} not expected in a bundle entry')
        out.setdefault(tag >> 3, []).append(v)
    return out


class TensorBundle:
    """Read-only view of one TensorFlow checkpoint prefix."""

    def __init__(self, prefix):
        self.prefix = prefix
        with open(prefix + '.index', 'rb') as f:
            buf = f.read()
        if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != TABLE_MAGIC:
            raise ValueError(f'{prefix}.index is not a TensorFlow checkpoint index')
        foot = _Cursor(buf, len(buf) - 48)
        foot.varint(); foot.varint()                       # metaindex handle
        ioff, isize = foot.varint(), foot.varint()
        self.entries = {}
        for _, handle in _table_block(buf, ioff, isize):
            h = _Cursor(handle)
            for key, val in _table_block(buf, h.varint(), h.varint()):
                if not key:
                    continue                               # header entry
                f = _proto_fields(val)
                shape = tuple(_proto_fields(d).get(1, [0])[0] for d in _proto_fields(f[2][0]).get(2, [])) if 2 in f else ()
                self.entries[key.decode()] = (f.get(1, [0])[0], shape, f.get(3, [0])[0], f.get(4, [0])[0], f.get(5, [0])[0])
        folder, base = os.path.split(prefix)
        self.nshards = len([f for f in os.listdir(folder or '.') if f.startswith(base + '.data-')])

    def keys(self):
        return list(self.entries)

    def tensor(self, key):
        dtype, shape, shard, off, size = self.entries[key]
        with open(f'{self.prefix}.data-{shard:05d}-of-{self.nshards:05d}', 'rb') as f:
            f.seek(off)
            raw = f.read(size)
        return np.frombuffer(raw, dtype=NP_DTYPE[dtype]).reshape(shape).copy()

    def keras_layers(self):
        """[{attr: array}] for layer_with_weights-0, -1, ... (attr = kernel | depthwise_kernel | bias | gamma | beta |
        moving_mean | moving_variance | cell/kernel | cell/recurrent_kernel | cell/bias)"""
        layers = {}
        for key, e in self.entries.items():
            if not key.endswith(SUFFIX) or not key.startswith('layer_with_weights-') or e[0] not in NP_DTYPE:
                continue
            head, attr = key[:-len(SUFFIX)].split('/', 1)
            layers.setdefault(int(head.split('-')[1]), {})[attr] = self.tensor(key)
        return [layers[i] for i in sorted(layers)]


# ---------------------------------------------------------------------------------------------- layer order
def _unit_names(arena_names):
    """tower unit prefixes in construction order, with their stride (a unit with a shortcut branch has stride 2)"""
    units, have = [], set(arena_names)
    for n in arena_names:
        if n.startswith('tower.s') and n.endswith('.pw1.w'):
            u = n[:-len('.pw1.w')]
            units.append((u, 2 if (u + '.scdw.w') in have else 1))
    return units


def keras_layer_sequence(model, arena_names):
    """[(arena prefix, kind)] in the order Keras numbers the weighted layers of `model`
    ('dynamics' | 'policy' | 'value'); kind in conv | dw | bn | dense | gru.  Layers at the same depth of the functional
    graph are numbered shortcut-before-branch (verified against the six shipped checkpoints)."""
    if model != 'dynamics':
        heads = ('alpha', 'beta', 'similarity', 'speed') if model == 'policy' else ('base', 'exp', 'speed', 'similarity')
        return [('bn1', 'bn'), ('d1', 'dense'), ('bn2', 'bn'), ('d2', 'dense')] + [(h, 'dense') for h in heads]
    seq = [('tower.stem', 'conv'), ('tower.stem', 'bn')]
    for u, stride in _unit_names(list(arena_names)):
        seq += [(u + '.pw1', 'conv'), (u + '.pw1', 'bn')]
        if stride == 2:
            seq += [(u + '.scdw', 'dw'), (u + '.dw', 'dw'), (u + '.scdw', 'bn'), (u + '.dw', 'bn'),
                    (u + '.scpw', 'conv'), (u + '.pw2', 'conv'), (u + '.scpw', 'bn'), (u + '.pw2', 'bn')]
        else:
            seq += [(u + '.dw', 'dw'), (u + '.dw', 'bn'), (u + '.pw2', 'conv'), (u + '.pw2', 'bn')]
    feats = ('road', 'vehicle', 'navigation')
    seq += [('tower.head', 'conv')] + [(f'feat.{f}.d1', 'dense') for f in feats] + [('tower.head', 'bn')]
    seq += [(f'feat.{f}.d1', 'bn') for f in feats] + [(f'feat.{f}.d2', 'dense') for f in feats]
    seq += [(f'feat.{f}.d2', 'bn') for f in feats]
    seq += [(f'gru.{g}', 'gru') for g in ('image',) + feats]
    seq += [('trunk.bn', 'bn'), ('trunk.dense', 'dense')]
    return seq


def bundle_to_arena_dict(prefix, model, arena_names):
    """the reference checkpoint at `prefix` -> {arena tensor name: np.ndarray} (trainable and moving statistics)"""
    layers = TensorBundle(prefix).keras_layers()
    seq = keras_layer_sequence(model, arena_names)
    if len(layers) != len(seq):
        raise ValueError(f'{prefix}: {len(layers)} weighted layers, expected {len(seq)} for the {model} model')
    out = {}
    for lay, (name, kind) in zip(layers, seq):
        if kind == 'bn':
            out[name + '.g'], out[name + '.be'] = lay['gamma'], lay['beta']
            out[name + '.mm'], out[name + '.mv'] = lay['moving_mean'], lay['moving_variance']
        elif kind == 'gru':
            out[name + '.k'], out[name + '.r'], out[name + '.b'] = lay['cell/kernel'], lay['cell/recurrent_kernel'], lay['cell/bias']
        else:
            k = lay['depthwise_kernel'] if kind == 'dw' else lay['kernel']
            if kind == 'dw':
                k = k.reshape(3, 3, k.shape[2])              # Keras [3,3,C,1]
            elif kind == 'conv' and k.ndim == 4 and k.shape[0] == 1:
                k = k.reshape(k.shape[2], k.shape[3])        # Keras [1,1,K,N]
            out[name + '.w'], out[name + '.b'] = k, lay['bias']
    return out


# ---------------------------------------------------------------------------------------------- arenas <-> files
def save_model(filepath, named_views):
    """`named_views`: iterable of (name, torch view) -> `<filepath>.npz`"""
    os.makedirs(os.path.dirname(filepath) or '.', exist_ok=True)
    np.savez(filepath + '.npz', **{n: v.detach().cpu().numpy() for n, v in named_views})


def load_model(filepath, model, named_views):
    """fills the (name, torch view) pairs from `<filepath>.npz` or from the reference's `<filepath>.index` bundle;
    raises FileNotFoundError / ValueError (shape mismatch) like Keras does"""
    import torch
    named_views = list(named_views)
    if os.path.exists(filepath + '.npz'):
        z = np.load(filepath + '.npz')
        src = {n: z[n] for n in z.files}
    elif os.path.exists(filepath + '.index'):
        src = bundle_to_arena_dict(filepath, model, [n for n, _ in named_views])
    else:
        raise FileNotFoundError(f'no checkpoint at {filepath} (.npz or TensorFlow .index)')
    for n, v in named_views:
        if n not in src:
            raise ValueError(f'{filepath}: variable {n} missing')
        a = np.asarray(src[n], dtype=np.float32)
        if tuple(a.shape) != tuple(v.shape):
            raise ValueError(f'{filepath}: {n} has shape {tuple(a.shape)}, expected {tuple(v.shape)}')
        v.copy_(torch.from_numpy(a))
