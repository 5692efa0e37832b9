"""Writes a TensorFlow tensor-bundle checkpoint (`<prefix>.index` + `<prefix>.data-00000-of-00001`) from numpy arrays.
TEST INFRASTRUCTURE: lets the checkpoint-import tests run where the reference's shipped `weights/stage-*` files are not
available (the GPU box) -- the golden weights in `ckpt_s5_curriculum.npz` are re-emitted in the reference's on-disk
format with Keras' variable naming, and the product's reader (cdra/checkpoint.py) has to bring them back.

Format (same as the files under /root/reference/weights): LevelDB table, no compression; data blocks of prefix-compressed
entries with a restart point every 16 keys; every block is followed by a 5-byte trailer (type 0 + crc32c, which readers of
the bundle index do not verify and which is left zero here); index block of (last key, BlockHandle); 48-byte footer."""
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
DT = {np.dtype('float32'): 1, np.dtype('float64'): 2, np.dtype('int32'): 3, np.dtype('int64'): 9}


def _vi(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(num, wt, payload):
    return _vi((num << 3) | wt) + payload


def _entry_proto(arr, offset):
    shape = b''.join(_field(2, 2, _vi(len(d)) + d) for d in (_field(1, 0, _vi(int(s))) for s in arr.shape))
    msg = _field(1, 0, _vi(DT[arr.dtype])) + _field(2, 2, _vi(len(shape)) + shape)
    msg += _field(4, 0, _vi(offset)) + _field(5, 0, _vi(arr.nbytes)) + _field(6, 5, struct.pack('<I', 0))
    return msg                                            # shard_id 0 is the proto default: omitted


def _block(items, interval=16):
    out, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def write_bundle(prefix, tensors, block_entries=40):
    """tensors: {checkpoint key: np.ndarray}"""
    keys = sorted(tensors)
    data, entries = bytearray(), []
    header = _field(1, 0, _vi(1)) + _field(3, 2, _vi(2) + _field(1, 0, _vi(1)))          # num_shards 1, version.producer 1
    entries.append((b'', header))
    for k in keys:
        a = np.ascontiguousarray(tensors[k])
        entries.append((k.encode(), _entry_proto(a, len(data))))
        data += a.tobytes()
    with open(f'{prefix}.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))
    out, index_items = bytearray(), []
    for i in range(0, len(entries), block_entries):
        chunk = entries[i:i + block_entries]
        blk = _block(chunk)
        index_items.append((chunk[-1][0], _vi(len(out)) + _vi(len(blk))))
        out += blk + b'\x00' * 5
    meta_off = len(out)
    meta = _block([])
    out += meta + b'\x00' * 5
    idx_off = len(out)
    idx = _block(index_items, interval=1)
    out += idx + b'\x00' * 5
    footer = _vi(meta_off) + _vi(len(meta)) + _vi(idx_off) + _vi(len(idx))
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC)
    out += footer
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))


def keras_checkpoint_tensors(params, order):
    """{oracle name: array} + [(prefix, kind)] in Keras layer order (oracle/ckpt.py) -> {checkpoint key: array} with the
    shapes Keras stores ([1,1,K,N] pointwise kernels, [3,3,C,1] depthwise kernels)"""
    out = {}
    sfx = '/.ATTRIBUTES/VARIABLE_VALUE'
    for i, (name, kind) in enumerate(order):
        base = f'layer_with_weights-{i}/'
        f = lambda n: np.asarray(params[n], dtype=np.float32)
        if kind == 'bn':
            out[base + 'gamma' + sfx], out[base + 'beta' + sfx] = f(name + '.g'), f(name + '.be')
            out[base + 'moving_mean' + sfx], out[base + 'moving_variance' + sfx] = f(name + '.mm'), f(name + '.mv')
        elif kind == 'gru':
            out[base + 'cell/kernel' + sfx], out[base + 'cell/recurrent_kernel' + sfx] = f(name + '.k'), f(name + '.r')
            out[base + 'cell/bias' + sfx] = f(name + '.b')
        else:
            w = f(name + '.w')
            if kind == 'dw':
                out[base + 'depthwise_kernel' + sfx] = w.reshape(3, 3, w.shape[2], 1)
            elif kind == 'conv':
                out[base + 'kernel' + sfx] = w.reshape(1, 1, *w.shape) if w.ndim == 2 else w
            else:
                out[base + 'kernel' + sfx] = w
            out[base + 'bias' + sfx] = f(name + '.b')
    # a few non-variable entries like the real files carry (object graph, save counter): the reader must skip them
    out['_CHECKPOINTABLE_OBJECT_GRAPH'] = np.zeros(3, dtype=np.int32)
    out['save_counter' + sfx] = np.asarray(1, dtype=np.int64)
    return out
