"""PyTorch-CPU restatement of the reference networks (oracle; test infrastructure).

Layout conventions: images NHWC like the reference; the time axis is explicit, inputs are
`[B, T, ...]` exactly as `CARLANetwork._get_input_layers` declares (core/networks.py:237-245).
Parameters are a dict name -> tensor following `oracle/spec.py`.  Gradients come from autograd.
Every function cites the reference lines it restates; [lib] marks Keras/TFP library semantics.
"""
import math
import torch
import torch.nn.functional as F

from . import spec

EPSILON = 1.1920928955078125e-07          # rl/utils.py:24-25 (np.finfo(float32).eps)


# ----------------------------------------------------------------------------- init
def glorot_uniform_(t, fan_in, fan_out, gen):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return t.uniform_(-lim, lim, generator=gen)


def _fans(shape):
    """Keras `_compute_fans` [lib]."""
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = 1
    for s in shape[:-2]:
        rf *= s
    return shape[-2] * rf, shape[-1] * rf


def init_params(params_spec, seed=42, dtype=torch.float32, zero_bias=()):
    """Keras-default initialisation restated: glorot-uniform kernels, glorot-uniform biases where the
    reference passes `bias_initializer='glorot_uniform'` (Dense layers and GRUs; conv biases are
    zeros [lib]), BN gamma=1 beta=0, moving mean 0 / variance 1.  Conv biases get small random values
    instead of zeros when `zero_bias` does not list them so that parity tests exercise the bias path."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape, kind in params_spec:
        t = torch.empty(shape, dtype=torch.float64)
        if kind == 'w':
            if name.endswith('.dw.w') or name.endswith('.scdw.w'):
                # depthwise kernel [3,3,C] == Keras [3,3,C,1]: fan_in = 9*C ... Keras uses (9*C, 9*1)
                fi, fo = 9 * shape[2], 9
            else:
                fi, fo = _fans(shape)
            glorot_uniform_(t, fi, fo, gen)
        elif kind == 'b':
            if len(shape) == 2:                       # GRU bias [2, 3u]
                glorot_uniform_(t, shape[0], shape[1], gen)
            else:
                glorot_uniform_(t, shape[0], shape[0], gen)
                if name.startswith('tower.'):
                    t.mul_(0.25)
        elif kind in ('g', 'mv'):
            t.fill_(1.0)
        else:
            t.zero_()
        out[name] = t.to(dtype)
    return out


def randomize_bn(params, seed=7):
    """Perturb BN gamma/beta/moving stats so tests do not run on the identity affine."""
    gen = torch.Generator().manual_seed(seed)
    for name, t in params.items():
        if name.endswith('.g'):
            t.copy_(1.0 + 0.2 * torch.randn(t.shape, generator=gen, dtype=torch.float64).to(t.dtype))
        elif name.endswith('.be'):
            t.copy_(0.2 * torch.randn(t.shape, generator=gen, dtype=torch.float64).to(t.dtype))
        elif name.endswith('.mm'):
            t.copy_(0.1 * torch.randn(t.shape, generator=gen, dtype=torch.float64).to(t.dtype))
        elif name.endswith('.mv'):
            t.copy_(0.5 + torch.rand(t.shape, generator=gen, dtype=torch.float64).to(t.dtype))
    return params


# ----------------------------------------------------------------------------- primitives
class BNState:
    """Collects what a training-mode BN call produces besides its output."""

    def __init__(self):
        self.updates = {}     # name -> list of (mean, var_for_moving) per call (time slice)

    def add(self, name, mean, var):
        self.updates.setdefault(name, []).append((mean.detach(), var.detach()))

    def apply_moving(self, params):
        """Keras: moving <- moving*momentum + batch*(1-momentum), once per call, in call order
        (core/architectures.py:44-57 => 4 sequential updates per forward) [lib]."""
        m = spec.BN_MOMENTUM
        new = {}
        for name, calls in self.updates.items():
            mm = params[name + '.mm'].clone()
            mv = params[name + '.mv'].clone()
            for mean, var in calls:
                # Keras assign_moving_average: variable -= (variable - value) * (1 - momentum) [lib]
                mm = mm - (mm - mean.to(mm.dtype)) * (1.0 - m)
                mv = mv - (mv - var.to(mv.dtype)) * (1.0 - m)
            new[name + '.mm'] = mm
            new[name + '.mv'] = mv
        return new


def batch_norm(x, p, name, training, bn_state=None):
    """Keras BatchNormalization(axis=-1), eps 1e-3, momentum 0.99 [lib].
    x: [..., C] (NHWC or [B, C]).  Training: biased batch variance normalises; the moving variance
    is fed the unbiased one for 4-D inputs (FusedBatchNormV3) and the biased one for 2-D inputs."""
    g, be = p[name + '.g'], p[name + '.be']
    if training:
        dims = tuple(range(x.dim() - 1))
        n = x.numel() // x.shape[-1]
        mean = x.mean(dim=dims)
        var = ((x - mean) ** 2).mean(dim=dims)
        if bn_state is not None:
            var_m = var * (n / max(n - 1, 1)) if x.dim() == 4 else var
            bn_state.add(name, mean, var_m)
    else:
        mean, var = p[name + '.mm'], p[name + '.mv']
    return (x - mean) * torch.rsqrt(var + spec.BN_EPS) * g + be


def relu6(x):
    return torch.clamp(x, 0.0, 6.0)


def swish6(x):
    """rl/utils.py:420-421."""
    return torch.minimum(x * torch.sigmoid(x), torch.full_like(x, 6.0))


def conv1x1(x, w, b):
    """Conv2D kernel 1 (core/architectures.py:130,134,140,170); x NHWC, w [Cin,Cout]."""
    return x @ w + b


def _pad_same_nchw(x, k, stride, value=0.0):
    h, w = x.shape[-2:]
    pt, pb = spec.same_pad(h, k, stride)
    pl, pr = spec.same_pad(w, k, stride)
    return F.pad(x, (pl, pr, pt, pb), value=value)


def depthwise3x3(x, w, b, stride):
    """DepthwiseConv2D 3x3 'same' (core/architectures.py:132,138); x NHWC, w [3,3,C]; TF SAME pad."""
    c = x.shape[-1]
    xc = _pad_same_nchw(x.permute(0, 3, 1, 2), 3, stride)
    wc = w.permute(2, 0, 1).unsqueeze(1)                   # [C,1,3,3]
    y = F.conv2d(xc, wc, bias=b, stride=stride, groups=c)
    return y.permute(0, 2, 3, 1)


def stem_conv(x, w, b):
    """Conv2D 3x3 stride 2 'valid' 3->24 (core/architectures.py:159); w [3,3,3,24] (HWIO)."""
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), bias=b, stride=2)
    return y.permute(0, 2, 3, 1)


def maxpool3x3s2(x):
    """MaxPooling2D(3, strides 2, 'same') (core/architectures.py:161); pads with -inf [lib]."""
    xc = _pad_same_nchw(x.permute(0, 3, 1, 2), 3, 2, value=float('-inf'))
    return F.max_pool2d(xc, 3, 2).permute(0, 2, 3, 1)


def channel_shuffle(x, groups=2):
    """core/architectures.py:109-118: reshape [..., C/g, g] -> transpose -> reshape
    => out[g*C/2 + i] = in[2i + g] (de-interleave)."""
    c = x.shape[-1]
    lead = x.shape[:-1]
    return x.reshape(*lead, c // groups, groups).transpose(-1, -2).reshape(*lead, c)


def shuffle_perm(c, groups=2):
    """Source index for every output channel of `channel_shuffle` (pure index form)."""
    return [(j % (c // groups)) * groups + j // (c // groups) for j in range(c)]


# ----------------------------------------------------------------------------- tower
def shufflenet_unit(x, p, name, stride, c, training, bn_state, taps=None):
    """core/architectures.py:120-145."""
    if stride == 1:
        half = x.shape[-1] // 2
        shortcut, y = x[..., :half], x[..., half:]
    else:
        shortcut, y = x, x

    def bn(t, n):
        return batch_norm(t, p, n, training, bn_state)

    y = conv1x1(y, p[name + '.pw1.w'], p[name + '.pw1.b'])
    if taps is not None:
        taps[name + '.pw1'] = y
    y = relu6(bn(y, name + '.pw1'))
    y = depthwise3x3(y, p[name + '.dw.w'], p[name + '.dw.b'], stride)
    if taps is not None:
        taps[name + '.dw'] = y
    y = bn(y, name + '.dw')
    y = conv1x1(y, p[name + '.pw2.w'], p[name + '.pw2.b'])
    if taps is not None:
        taps[name + '.pw2'] = y
    y = relu6(bn(y, name + '.pw2'))
    if stride == 2:
        s = depthwise3x3(shortcut, p[name + '.scdw.w'], p[name + '.scdw.b'], 2)
        if taps is not None:
            taps[name + '.scdw'] = s
        s = bn(s, name + '.scdw')
        s = conv1x1(s, p[name + '.scpw.w'], p[name + '.scpw.b'])
        if taps is not None:
            taps[name + '.scpw'] = s
        shortcut = relu6(bn(s, name + '.scpw'))
    out = channel_shuffle(torch.cat([shortcut, y], dim=-1))
    if taps is not None:
        taps[name + '.out'] = out
    return out


def tower_slice(x, p, training, bn_state, taps=None):
    """One time slice through the shared-weight tower (core/architectures.py:153-173); x [B,H,W,3]."""
    y = stem_conv(x, p['tower.stem.w'], p['tower.stem.b'])
    if taps is not None:
        taps['tower.stem'] = y
    y = relu6(batch_norm(y, p, 'tower.stem', training, bn_state))
    y = maxpool3x3s2(y)
    if taps is not None:
        taps['tower.pool'] = y
    for name, stride, cin, c in spec.tower_units():
        y = shufflenet_unit(y, p, name, stride, c, training, bn_state, taps)
    y = conv1x1(y, p['tower.head.w'], p['tower.head.b'])
    if taps is not None:
        taps['tower.head'] = y
    y = relu6(batch_norm(y, p, 'tower.head', training, bn_state))
    return y.mean(dim=(1, 2))                             # GlobalAveragePooling2D


def feature_net_slice(x, p, fname, training, bn_state):
    """core/architectures.py:9-27 with units=16, num_layers=2, activation=relu6, normalization=None:
    Dense(relu6) -> BN -> Dense(relu6) -> BN."""
    for d in ('d1', 'd2'):
        n = f'feat.{fname}.{d}'
        x = relu6(x @ p[n + '.w'] + p[n + '.b'])
        x = batch_norm(x, p, n, training, bn_state)
    return x


def gru(x_seq, k, r, b):
    """Keras GRU, reset_after=True, gate order [z|r|h], h0=0, returns last state
    (core/networks.py:47-50) [lib].  x_seq: list of [B, D]."""
    units = r.shape[0]
    h = torch.zeros(x_seq[0].shape[0], units, dtype=x_seq[0].dtype)
    for x in x_seq:
        xp = x @ k + b[0]
        hp = h @ r + b[1]
        xz, xr, xh = xp.split(units, dim=1)
        hz, hr, hh = hp.split(units, dim=1)
        z = torch.sigmoid(xz + hz)
        rr = torch.sigmoid(xr + hr)
        hc = torch.tanh(xh + rr * hh)
        h = z * h + (1.0 - z) * hc
    return h


def dynamics_forward(p, obs, training=True, bn_state=None, taps=None):
    """core/networks.py:37-56.  obs: dict state_image [B,T,H,W,3] in [0,1], state_road [B,T,9],
    state_vehicle [B,T,4], state_navigation [B,T,5].  Returns dynamics_out [B,512]."""
    T = obs['state_image'].shape[1]
    img = [tower_slice(obs['state_image'][:, t], p, training, bn_state,
                       taps if (taps is not None and t == 0) else None) for t in range(T)]
    if taps is not None:
        taps['tower.gap'] = torch.stack(img, dim=0)       # [T,B,768]
    feats = {}
    for fname, _ in spec.FEATURES:
        feats[fname] = [feature_net_slice(obs['state_' + fname][:, t], p, fname, training, bn_state)
                        for t in range(T)]
    outs = [gru(img, p['gru.image.k'], p['gru.image.r'], p['gru.image.b'])]
    for fname, _ in spec.FEATURES:
        outs.append(gru(feats[fname], p[f'gru.{fname}.k'], p[f'gru.{fname}.r'], p[f'gru.{fname}.b']))
    x = torch.cat(outs, dim=1)                             # dynamics_in [B,352]
    if taps is not None:
        taps['dynamics_in'] = x
    x = batch_norm(x, p, 'trunk.bn', training, bn_state)   # linear_combination, networks.py:24-30
    return x @ p['trunk.dense.w'] + p['trunk.dense.b']


# ----------------------------------------------------------------------------- heads
def control_branch(x, p, training, bn_state=None):
    """core/networks.py:59-66: 2 x [BN -> Dense(320, swish6)]."""
    for i in (1, 2):
        x = batch_norm(x, p, f'bn{i}', training, bn_state)
        x = swish6(x @ p[f'd{i}.w'] + p[f'd{i}.b'])
    return x


def softplus_c(x, c=1.0 + 1e-2):
    """rl/utils.py:411-416 with value 1.01 (core/networks.py:133-134)."""
    return F.softplus(x) + c


def beta_log_prob(a, b, x):
    """tfp.distributions.Beta(concentration1=a, concentration0=b).log_prob [lib]."""
    lbeta = torch.lgamma(a) + torch.lgamma(b) - torch.lgamma(a + b)
    return (a - 1.0) * torch.log(x) + (b - 1.0) * torch.log1p(-x) - lbeta


def beta_entropy(a, b):
    lbeta = torch.lgamma(a) + torch.lgamma(b) - torch.lgamma(a + b)
    return (lbeta - (a - 1.0) * torch.digamma(a) - (b - 1.0) * torch.digamma(b)
            + (a + b - 2.0) * torch.digamma(a + b))


def beta_sample_reparameterized(alpha, beta, generator=None):
    """tfp.distributions.Beta._sample_n [lib] (tensorflow-probability 0.11): x = g1 / (g1 + g2) with g1 ~ Gamma(alpha),
    g2 ~ Gamma(beta); the distribution is FULLY_REPARAMETERIZED, the gamma draws carry implicit-reparameterisation
    gradients dg/dalpha = -(dF/dalpha) / pdf (tf.random.gamma's RandomGammaGrad; torch exposes the same quantity as
    torch._standard_gamma_grad).  Returns (sample, jac [..., 2] = (dx/dalpha, dx/dbeta)) as constants."""
    with torch.no_grad():
        a, b = alpha.detach(), beta.detach()
        g1 = torch._standard_gamma(a, generator=generator) if generator is not None else torch._standard_gamma(a)
        g2 = torch._standard_gamma(b, generator=generator) if generator is not None else torch._standard_gamma(b)
        s = g1 + g2
        x = g1 / s
        dg1, dg2 = torch._standard_gamma_grad(a, g1), torch._standard_gamma_grad(b, g2)
        jac = torch.stack([g2 / (s * s) * dg1, -g1 / (s * s) * dg2], dim=-1)
    return x, jac


def policy_forward(p, x512, actions_eval, training=True, bn_state=None, actions_jac=None):
    """PolicyNetwork.call (core/networks.py:96-110).  `actions_eval` stands for the fresh sample of the new policy the
    reference draws (:97-100; decision D2); it is clipped like `_clip_actions` (:139-144).  With `actions_jac`
    [B,2,2] = (d a / d alpha, d a / d beta) the action is a reparameterised sample and the log-prob is differentiated through
    it, as TFP does; without it the action is a constant."""
    h = control_branch(x512, p, training, bn_state)
    alpha = softplus_c(h @ p['alpha.w'] + p['alpha.b'])
    beta = softplus_c(h @ p['beta.w'] + p['beta.b'])
    sim = torch.tanh(h @ p['similarity.w'] + p['similarity.b'])
    speed = 2.0 * torch.sigmoid(h @ p['speed.w'] + p['speed.b'])
    if actions_jac is not None:          # value of the sample unchanged, first-order dependence on (alpha, beta) attached
        actions_eval = (actions_eval + actions_jac[..., 0] * (alpha - alpha.detach()) + actions_jac[..., 1] * (beta - beta.detach()))
    a = actions_eval.clamp(EPSILON, 1.0 - EPSILON)      # tf.clip_by_value: gradient passes where not clipped
    logp = beta_log_prob(alpha, beta, a)
    ent = beta_entropy(alpha, beta)
    mean = alpha / (alpha + beta)
    std = torch.sqrt(alpha * beta / ((alpha + beta) ** 2 * (alpha + beta + 1.0)))
    return dict(alpha=alpha, beta=beta, log_prob=logp, entropy=ent, mean=mean.detach(), std=std.detach(),
                speed=speed, similarity=sim)


def value_forward(p, x512, training=True, bn_state=None, exp_scale=6.0):
    """CARLANetwork.value_branch / value_head (core/networks.py:255-275)."""
    h = control_branch(x512, p, training, bn_state)
    base = torch.tanh(h @ p['base.w'] + p['base.b'])
    exp = exp_scale * torch.sigmoid(h @ p['exp.w'] + p['exp.b'])
    speed = 2.0 * torch.sigmoid(h @ p['speed.w'] + p['speed.b'])
    sim = torch.tanh(h @ p['similarity.w'] + p['similarity.b'])
    return dict(value=torch.cat([base, exp], dim=1), speed=speed, similarity=sim)
