"""Minimal reader for TensorFlow "tensor bundle" checkpoints (the format of the reference's shipped
`weights/stage-*/{dynamics_model,policy_net,value_net}.{index,data-*}`).

`.index` is an uncompressed LevelDB-style SSTable: data blocks of prefix-compressed (key, value)
entries, an index block, and a 48-byte footer.  Values are `BundleEntryProto` messages
{1: dtype, 2: TensorShapeProto, 3: shard_id, 4: offset, 5: size, 6: crc32c}.
Tensors are raw little-endian in `<prefix>.data-<shard:05d>-of-<n:05d>`.
Pure python + numpy; test infrastructure (also used by the checkpoint-import "next" row, §8f-3).
"""
import os
import struct
import numpy as np

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _block_entries(buf, offset, size):
    block = buf[offset:offset + size]
    n_restarts = struct.unpack('<I', block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _parse_proto(buf):
    """-> dict field -> list of (wire_type, value)."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = struct.unpack('<I', buf[pos:pos + 4])[0]
            pos += 4
        elif wt == 1:
            val = struct.unpack('<Q', buf[pos:pos + 8])[0]
            pos += 8
        else:
            raise ValueError(f'unsupported wire type {wt}')
        out.setdefault(field, []).append(val)
    return out


def _shape(buf):
    dims = []
    for d in _parse_proto(buf).get(2, []):
        dims.append(_parse_proto(d).get(1, [0])[0])
    return tuple(dims)


def read_index(prefix):
    """-> {variable key: dict(dtype, shape, shard, offset, size)} (header entry '' skipped)."""
    with open(prefix + '.index', 'rb') as f:
        buf = f.read()
    footer = buf[-48:]
    assert footer[-8:] == struct.pack('<Q', 0xdb4775248b80fb57), 'bad sstable magic'
    pos = 0
    _, pos = _varint(footer, pos)
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _block_entries(buf, idx_off, idx_size):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, value in _block_entries(buf, off, size):
            if key == b'':
                continue
            msg = _parse_proto(value)
            entries[key.decode()] = dict(dtype=msg.get(1, [0])[0], shape=_shape(msg[2][0]) if 2 in msg else (),
                                         shard=msg.get(3, [0])[0], offset=msg.get(4, [0])[0],
                                         size=msg.get(5, [0])[0])
    return entries


def read_tensors(prefix, only_variables=True):
    """-> {key: np.ndarray} for every float/int tensor entry."""
    entries = read_index(prefix)
    folder, base = os.path.split(prefix)
    shards = sorted(f for f in os.listdir(folder or '.') if f.startswith(base + '.data-'))
    n = len(shards)
    out = {}
    for key, e in entries.items():
        if only_variables and not key.endswith('.ATTRIBUTES/VARIABLE_VALUE'):
            continue
        if e['dtype'] not in _DTYPES:
            continue
        path = f'{prefix}.data-{e["shard"]:05d}-of-{n:05d}'
        with open(path, 'rb') as f:
            f.seek(e['offset'])
            raw = f.read(e['size'])
        out[key] = np.frombuffer(raw, dtype=_DTYPES[e['dtype']]).reshape(e['shape']).copy()
    return out


def layer_variables(prefix):
    """Group variables by Keras `layer_with_weights-N` -> {N: {attr: array}} (attr = kernel, bias,
    gamma, beta, moving_mean, moving_variance, depthwise_kernel, cell/kernel, ...)."""
    out = {}
    for key, arr in read_tensors(prefix).items():
        parts = key.split('/')
        if not parts[0].startswith('layer_with_weights-'):
            continue
        idx = int(parts[0].split('-')[1])
        attr = '/'.join(parts[1:-2])
        out.setdefault(idx, {})[attr] = arr
    return out
