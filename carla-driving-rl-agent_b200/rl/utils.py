"""Numeric subset of the reference's rl/utils.py needed by the PPO-update path, on PyTorch
tensors + libcdra.  Plotting / trace IO / gym helpers that the hot path never calls are not mirrored."""
import os
import random
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from rl import spaces
from rl.parameters import DynamicParameter

NP_EPS = np.finfo(np.float32).eps            # rl/utils.py:24
EPSILON = float(NP_EPS)                      # rl/utils.py:25
OPTIMIZERS = ('adam',)                       # the only optimiser the CUDA path implements (rl/utils.py:29-46 lists 8)


def get_optimizer_by_name(name: str, *args, **kwargs):
    """rl/utils.py:39-46: unknown names raise ValueError; here everything but Adam is 'unknown'."""
    if name.lower() not in OPTIMIZERS:
        raise ValueError(f'Cannot find optimizer {name}. Select one of {OPTIMIZERS}.')
    print(f'Optimizer: {name}.')
    return dict(name=name.lower(), args=args, kwargs=kwargs)


def makedir(*args: str) -> str:
    path = os.path.join(*args)
    os.makedirs(path, exist_ok=True)
    return path


def to_float(x):
    return x.float() if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float32)


def to_tensor(x, expand_axis=0):
    """rl/utils.py `to_tensor`: python / numpy structures -> float tensors with a leading batch axis."""
    if isinstance(x, dict):
        return {k: to_tensor(v, expand_axis) for k, v in x.items()}
    t = torch.as_tensor(np.asarray(x), dtype=torch.float32) if not isinstance(x, torch.Tensor) else x.float()
    return t.unsqueeze(expand_axis) if expand_axis is not None else t


def space_to_flat_spec(space, name: str) -> Dict[str, tuple]:
    """rl/utils.py:212-247."""
    kind = spaces.kind(space)
    spec = dict()
    if kind == 'discrete':
        spec[name] = (space.n,)
    elif kind == 'multidiscrete':
        spec[name] = space.nvec.shape
    elif kind == 'box':
        spec[name] = tuple(space.shape)
    else:
        for key, value in space.spaces.items():
            for k, v in space_to_flat_spec(value, f'{name}_{key}').items():
                spec[k] = v
    return spec


def clip(value, min_value, max_value):
    return min(max_value, max(value, min_value))


def decompose_number(num: float) -> Tuple[float, float]:
    """rl/utils.py:140-151 (fp32 arithmetic like the reference's tf.map_fn path)."""
    num = np.float32(num)
    exponent = 0
    while abs(num) > np.float32(1.0):
        num = np.float32(num / np.float32(10.0))
        exponent += 1
    return float(num), float(exponent)


def polyak_averaging(arena_flat: torch.Tensor, old_flat: torch.Tensor, alpha=0.99):
    """rl/utils.py:105-117: w = alpha * w_new + (1 - alpha) * w_old (in place on the flat arena)."""
    arena_flat.mul_(alpha).add_(old_flat, alpha=1.0 - alpha)


def index_batches(n: int, batch_size: int, shuffle_batches=False, seed=None, drop_remainder=False, num_shards=1, skip=0,
                  shuffle=False) -> List[np.ndarray]:
    """Index form of `data_to_batches` (rl/utils.py:365-393): tf.data `from_tensor_slices -> skip -> shuffle(buffer =
    batch_size) -> shard/concatenate -> batch(drop_remainder) -> shuffle(buffer = batch_size)`; returns the list of
    row-index arrays the CUDA gather kernel consumes (prefetch has no numerical meaning)."""
    rng = np.random.RandomState(seed if seed is not None else random.randint(0, 2 ** 31 - 1))
    idx = list(range(skip, n))

    def buffered_shuffle(items, buffer_size):
        """tf.data `shuffle(buffer_size)`: the buffer fills up, every further item pushes out a uniformly drawn one, the rest
        drains uniformly.  O(1) per item (swap-remove: the order inside the buffer has no meaning) with the random draws
        taken in bulk -- the update starts with this call while the GPU waits."""
        items = list(items)
        n, B = len(items), int(buffer_size)
        buf = items[:B]
        out = []
        if n > B:
            for x, j in zip(items[B:], rng.randint(0, B + 1, size=n - B).tolist()):
                buf.append(x)
                out.append(buf[j]); buf[j] = buf[-1]; buf.pop()
        m = len(buf)
        if m:
            sizes = np.arange(m, 0, -1)
            for j in np.minimum((rng.random_sample(m) * sizes).astype(np.int64), sizes - 1).tolist():
                out.append(buf[j]); buf[j] = buf[-1]; buf.pop()
        return out

    if shuffle:
        idx = buffered_shuffle(idx, batch_size)
    if num_shards > 1:
        idx = [i for s in range(num_shards) for i in idx[s::num_shards]]
    batches = [np.asarray(idx[i:i + batch_size], dtype=np.int64) for i in range(0, len(idx), batch_size)]
    if drop_remainder:
        batches = [b for b in batches if len(b) == batch_size]
    if shuffle_batches:
        batches = buffered_shuffle(batches, batch_size)
    return batches


class Summary:
    """utils.Summary (rl/utils.py:577-673): `agent.log(**kwargs)` collects values per key, `write_summaries()` emits one
    TensorBoard scalar per collected value (histograms for 'weight-' / 'bias-' keys, images for 'image_' keys) under
    `<summary_dir>/<name>/<timestamp>` and clears the lists; `mode='log'` only collects, `mode=None` disables.

    Values that live on the GPU (losses, ratios, per-tensor gradient norms written by the kernels) are kept as device
    tensors and reduced / copied to the host ONCE per `write_summaries()` -- logging never synchronises the update loop
    (the reference's eager `float(x)` per scalar would cost one host round trip per value per SGD step)."""

    def __init__(self, mode='summary', name=None, summary_dir='logs', keys: Optional[List[str]] = None):
        self.stats: Dict[str, dict] = {}
        self.allowed_keys = {k: True for k in keys} if isinstance(keys, (list, tuple, set)) else None
        self.should_log = mode in ('summary', 'log')
        self.use_summary = mode == 'summary'
        self.mode = mode
        self.name = name
        self.summary_dir = None
        self._writer = None
        self.last: Dict[str, float] = {}
        self.last_d2h_bytes = 0
        if self.use_summary:
            import datetime
            self.summary_dir = os.path.join(summary_dir, str(name), datetime.datetime.now().strftime('%Y%m%d-%H%M%S'))

    def should_log_key(self, key: str) -> bool:
        return True if self.allowed_keys is None else key in self.allowed_keys

    def log(self, **kwargs):
        if not self.should_log:
            return
        for key, value in kwargs.items():
            if not self.should_log_key(key):
                continue
            entry = self.stats.setdefault(key, dict(step=0, list=[]))
            if isinstance(value, np.ndarray):
                value = torch.from_numpy(np.atleast_1d(value))
            if isinstance(value, torch.Tensor):
                # a tensor with several elements extends the list row by row (tf: `list.extend(tensor)`); the snapshot is a
                # device-side copy, so buffers the kernels overwrite every step (loss scalars) can be logged as views
                v = value.detach()
                entry['list'].append(v.reshape(v.shape[0], -1).clone() if v.dim() >= 1 and v.numel() > 1 else v.reshape(1, 1).clone())
            elif isinstance(value, (list, tuple)) and len(value) and isinstance(value[0], torch.Tensor):
                entry['list'].append(torch.stack([x.detach().float().reshape(-1).mean() for x in value]).reshape(-1, 1))
            elif hasattr(value, '__iter__'):
                entry['list'].extend(np.mean(np.asarray(x, dtype=np.float64)) for x in value)
            else:
                entry['list'].append(float(value))

    def _collect(self):
        """{key: [float per logged value]}: every device tensor is reduced to per-row means on its device, all of them
        travel to the host in ONE copy"""
        dev_rows, slots = [], []
        out = {}
        for key, entry in self.stats.items():
            vals = []
            for item in entry['list']:
                if isinstance(item, torch.Tensor):
                    m = item.float().mean(dim=1)
                    if m.device.type == 'cpu':
                        vals.extend(m.tolist())
                    else:
                        slots.append((key, len(vals), m.numel()))
                        vals.extend([None] * m.numel())
                        dev_rows.append(m)
                else:
                    vals.append(float(item))
            out[key] = vals
        self.last_d2h_bytes = 0
        if dev_rows:
            packed = torch.cat(dev_rows)
            self.last_d2h_bytes = packed.numel() * 4
            host = packed.cpu().tolist()
            i = 0
            for key, pos, n in slots:
                out[key][pos:pos + n] = host[i:i + n]
                i += n
        return out

    def write_summaries(self):
        data = self._collect()
        self.last = {k: v[-1] for k, v in data.items() if v}
        if self.use_summary:
            if self._writer is None:
                from torch.utils.tensorboard import SummaryWriter
                self._writer = SummaryWriter(log_dir=self.summary_dir)
            for name, values in data.items():
                step = self.stats[name]['step']
                if 'weight-' in name or 'bias-' in name:
                    self._writer.add_histogram(name, np.asarray(values), global_step=step)
                else:
                    for i, v in enumerate(values):
                        self._writer.add_scalar(name, v, global_step=step + i)
            self._writer.flush()
        for name, values in data.items():
            self.stats[name]['step'] += len(values)
            self.stats[name]['list'].clear()

    def close(self):
        if self._writer is not None:
            self._writer.close()
            self._writer = None
