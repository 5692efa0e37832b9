"""Parameter inventory of the three reference models (oracle side; test infrastructure).

Follows the layer construction order of
  core/architectures.py:9-27   (feature_net)
  core/architectures.py:30-173 (shufflenet_v2)
  core/networks.py:37-66       (dynamics_layers, control_branch)
  core/networks.py:115-137     (PolicyNetwork.policy_branch / get_distribution_layer)
  core/networks.py:255-275     (CARLANetwork.value_branch / value_head)
with the default sizes of core/carla_agent.py:61-68.

Every entry is (name, shape, kind) with kind in
  'w'  kernel           'b' bias           'g' BN gamma      'be' BN beta     (trainable)
  'mm' BN moving mean   'mv' BN moving variance                               (state)
The flat "arena" order used by the CUDA library (csrc/plan.cpp) is exactly the order of
`dynamics_params()` / `head_params()` filtered by trainable / state; tests assert both agree.
"""
from collections import OrderedDict

STAGE_CHANNELS = (116, 232, 464)      # g = 1.0, core/architectures.py:34
STAGE_BLOCKS = (4, 8, 4)              # core/architectures.py:165-167
STEM_CHANNELS = 24                    # core/architectures.py:159
LAST_CHANNELS = 768                   # core/carla_agent.py:66
FEATURES = (('road', 9), ('vehicle', 4), ('navigation', 5))   # core/carla_env.py:20-27 sizes
FEAT_UNITS = 16
GRU_UNITS = (('image', LAST_CHANNELS, 256), ('road', 16, 32), ('vehicle', 16, 32), ('navigation', 16, 32))
TRUNK_IN = 256 + 32 * 3               # 352
TRUNK_UNITS = 512
HEAD_UNITS = 320
NUM_ACTIONS = 2
TIME_HORIZON = 4
BN_EPS = 1e-3
BN_MOMENTUM = 0.99


def _bn(prefix, c, out):
    out.append((prefix + '.g', (c,), 'g'))
    out.append((prefix + '.be', (c,), 'be'))
    out.append((prefix + '.mm', (c,), 'mm'))
    out.append((prefix + '.mv', (c,), 'mv'))


def _conv_bn(prefix, wshape, c, out):
    out.append((prefix + '.w', tuple(wshape), 'w'))
    out.append((prefix + '.b', (c,), 'b'))
    _bn(prefix, c, out)


def tower_units():
    """[(name, stride, cin, c)] for the 16 shufflenet units (core/architectures.py:147-167)."""
    units = []
    cin = STEM_CHANNELS
    for s, (c, nb) in enumerate(zip(STAGE_CHANNELS, STAGE_BLOCKS), start=1):
        for u in range(nb):
            units.append((f'tower.s{s}.u{u}', 2 if u == 0 else 1, cin, c))
            cin = c
    return units


def dynamics_params():
    out = []
    _conv_bn('tower.stem', (3, 3, 3, STEM_CHANNELS), STEM_CHANNELS, out)
    for name, stride, cin, c in tower_units():
        half = c // 2
        if stride == 2:
            sc = cin                       # shortcut_channels, core/architectures.py:127
            kin = cin
        else:
            sc = cin // 2
            kin = cin // 2
        _conv_bn(name + '.pw1', (kin, half), half, out)              # :130
        _conv_bn(name + '.dw', (3, 3, half), half, out)              # :132
        _conv_bn(name + '.pw2', (half, c - sc), c - sc, out)         # :134
        if stride == 2:
            _conv_bn(name + '.scdw', (3, 3, sc), sc, out)            # :138
            _conv_bn(name + '.scpw', (sc, sc), sc, out)              # :140
    _conv_bn('tower.head', (STAGE_CHANNELS[-1], LAST_CHANNELS), LAST_CHANNELS, out)   # :170
    for fname, d in FEATURES:
        _conv_bn(f'feat.{fname}.d1', (d, FEAT_UNITS), FEAT_UNITS, out)               # :20-25
        _conv_bn(f'feat.{fname}.d2', (FEAT_UNITS, FEAT_UNITS), FEAT_UNITS, out)
    for gname, din, units in GRU_UNITS:                                               # networks.py:47-50
        out.append((f'gru.{gname}.k', (din, 3 * units), 'w'))
        out.append((f'gru.{gname}.r', (units, 3 * units), 'w'))
        out.append((f'gru.{gname}.b', (2, 3 * units), 'b'))
    _bn('trunk.bn', TRUNK_IN, out)                                                    # networks.py:24-30
    out.append(('trunk.dense.w', (TRUNK_IN, TRUNK_UNITS), 'w'))
    out.append(('trunk.dense.b', (TRUNK_UNITS,), 'b'))
    return out


def head_params(kind):
    """kind = 'policy' | 'value' (core/networks.py:59-66,115-137,255-275)."""
    out = []
    _bn('bn1', TRUNK_UNITS, out)
    out.append(('d1.w', (TRUNK_UNITS, HEAD_UNITS), 'w'))
    out.append(('d1.b', (HEAD_UNITS,), 'b'))
    _bn('bn2', HEAD_UNITS, out)
    out.append(('d2.w', (HEAD_UNITS, HEAD_UNITS), 'w'))
    out.append(('d2.b', (HEAD_UNITS,), 'b'))
    if kind == 'policy':
        heads = (('alpha', NUM_ACTIONS), ('beta', NUM_ACTIONS), ('similarity', 1), ('speed', 1))
    else:
        heads = (('base', 1), ('exp', 1), ('speed', 1), ('similarity', 1))
    for hname, n in heads:
        out.append((f'{hname}.w', (HEAD_UNITS, n), 'w'))
        out.append((f'{hname}.b', (n,), 'b'))
    return out


def numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


def split_layout(params):
    """-> (trainable OrderedDict name->(offset, shape), n_trainable, state OrderedDict, n_state)."""
    tr, st = OrderedDict(), OrderedDict()
    ot = os_ = 0
    for name, shape, kind in params:
        if kind in ('mm', 'mv'):
            st[name] = (os_, shape)
            os_ += numel(shape)
        else:
            tr[name] = (ot, shape)
            ot += numel(shape)
    return tr, ot, st, os_


def conv_out_same(n, stride):
    return -(-n // stride)


def same_pad(n, k, stride):
    """TF 'SAME' padding (before, after); the extra cell goes after."""
    out = conv_out_same(n, stride)
    total = max((out - 1) * stride + k - n, 0)
    return total // 2, total - total // 2


def spatial_sizes(h, w):
    """stem(valid 3x3 s2) -> pool -> stage1..3 output sizes."""
    sh, sw = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    sizes = [(sh, sw)]
    ph, pw = conv_out_same(sh, 2), conv_out_same(sw, 2)
    sizes.append((ph, pw))
    for _ in range(3):
        ph, pw = conv_out_same(ph, 2), conv_out_same(pw, 2)
        sizes.append((ph, pw))
    return sizes
