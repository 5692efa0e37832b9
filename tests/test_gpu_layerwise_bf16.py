"""bf16 perf-mode kernels (the BENCHED path: v2_*.cuh incl. the tcgen05 kernels) against the oracle, ONE LAYER AT A
TIME on identical inputs: after a training step the library's own stored tensors of every layer (raw conv outputs,
BatchNorm tables, stored gradients; read back through cdra_debug_export in the reference's logical layout) are fed to
oracle/bf16_layers.py, which restates the layer with bf16 rounding at exactly the library's storage points.  Because the
inputs are identical, the 50-layer chaos of a training-mode BatchNorm tower is out of the picture and every kernel --
stem / pool / pointwise forward (mma.sync + tcgen05) / depthwise forward / GAP, and backward pw_dgrad / pw_wgrad(_tc) /
dw_bwd / gap_bwd / stem_bwd -- is held to bf16 resolution: forward taps <= 2e-3, stored gradients and parameter
gradients <= 1e-2 relative L2 (core/architectures.py:120-173, core/carla_agent.py:351-373)."""
import os

import pytest
import torch

from oracle import bf16_layers as L
from oracle import model, ppo, spec
from tests import common as C

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason='needs a CUDA device')]

H, W = 90, 120
FWD_TOL, GRAD_TOL, PGRAD_TOL, TAB_TOL = 2e-3, 1e-2, 1e-2, 2e-4


def _engine(B, h=H, w=W):
    from cdra.engine import Engine
    return Engine(B, h, w, dtype='bf16', image_u8=True, device='cuda')


def _dev(d):
    return {k: v.cuda() for k, v in d.items()}


class Probe:
    """reads the library's stored tensors / tables / gradients of one engine (all in the logical layout, fp64 on the host)"""

    def __init__(self, eng, dyn):
        self.eng, self.dyn = eng, dyn
        self.g = {k: v.double().cpu() for k, v in eng.dyn.to_dict(eng.g_dyn).items()}
        self.cache = {}

    def t(self, name):
        if name not in self.cache:
            self.cache[name] = self.eng.tensor(name).double().cpu()
        return self.cache[name]

    def tables(self, name):
        """(scale, shift, mean, inv) [4][C] as the library holds them"""
        a, b = self.t('aff:' + name).squeeze(-1), self.t('bnp:' + name).squeeze(-1)
        return a[..., 0], a[..., 1], b[..., 0], b[..., 1]

    def bsum(self, name):
        s = self.t('bsum:' + name).squeeze(-1)
        return s[..., 0], s[..., 1]

    def p(self, name):
        return self.dyn[name].double()


def _unit_layout(name, stride, cin, c):
    half = c // 2
    sc = cin if stride == 2 else cin // 2
    return half, sc


def _check_tables(rep, pr, name, raw, gname, chans=None):
    """the library's BatchNorm tables of a stored tensor against the statistics of the stored values"""
    g, be = pr.p(gname + '.g'), pr.p(gname + '.be')
    sc, sh, mu, inv = L.bn_tables(raw, g, be)
    got = pr.tables(name)
    if chans is not None:
        got = tuple(x[:, chans] for x in got)
    for nm, a, b in zip(('scale', 'shift', 'mean', 'inv'), got, (sc, sh, mu, inv)):
        e = ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
        rep.setdefault('tables', {})[f'{name}:{nm}'] = e


def _layerwise(eng, dyn, obs):
    """-> report {category: {name: error}} for every layer of the tower"""
    pr = Probe(eng, dyn)
    rep = {'fwd': {}, 'grad': {}, 'pgrad': {}, 'exact': {}, 'sums': {}}
    B = eng.B
    img = obs['state_image'].cpu()                                   # [B,4,H,W,3] u8
    frames = img.permute(1, 0, 2, 3, 4).reshape(4 * B, img.shape[2], img.shape[3], 3).double()

    def rel(a, b):
        return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()

    # ------------------------------------------------------------------ stem + pool (core/architectures.py:159-161)
    w_st = L.bf16r(L.f32r(pr.p('tower.stem.w') * L.f32r(torch.tensor(1.0 / 255.0, dtype=torch.float64))))
    stem_ref = L.bf16r(model.stem_conv(frames, w_st, pr.p('tower.stem.b')))
    stem = pr.t('tower.stem')
    rep['fwd']['tower.stem'] = rel(stem, stem_ref)
    _check_tables(rep, pr, 'tower.stem', stem, 'tower.stem')
    s_sc, s_sh, s_mu, s_inv = pr.tables('tower.stem')
    a_stem = L.activate(stem, s_sc, s_sh, True)
    pool_ref, pidx, padded_hw = L.maxpool_first(a_stem)
    pool = pr.t('tower.pool')
    rep['exact']['tower.pool'] = float((pool != pool_ref).sum().item())

    # ------------------------------------------------------------------ units
    x_raw, x_tabs, x_clamp, x_name = pool, None, False, 'tower.pool'
    units = spec.tower_units()
    for ui, (name, stride, cin, c) in enumerate(units):
        half, sc = _unit_layout(name, stride, cin, c)
        Cx = x_raw.shape[-1]
        if x_tabs is None:
            ones = torch.ones(4, Cx, dtype=torch.float64)
            x_tabs = (ones, torch.zeros_like(ones), torch.zeros_like(ones), ones)
        a_x = L.activate(x_raw, x_tabs[0], x_tabs[1], x_clamp) if x_clamp else x_raw.clone()
        xin = a_x if stride == 2 else a_x[..., Cx // 2:]
        # pw1
        r1 = pr.t(name + '.pw1')
        rep['fwd'][name + '.pw1'] = rel(r1, L.pw_forward(xin, pr.p(name + '.pw1.w'), pr.p(name + '.pw1.b')))
        _check_tables(rep, pr, name + '.pw1', r1, name + '.pw1')
        t1 = pr.tables(name + '.pw1')
        a1 = L.activate(r1, t1[0], t1[1], True)
        # dw
        r2 = pr.t(name + '.dw')
        rep['fwd'][name + '.dw'] = rel(r2, L.dw_forward(a1, pr.p(name + '.dw.w'), pr.p(name + '.dw.b'), stride))
        _check_tables(rep, pr, name + '.dw', r2, name + '.dw')
        t2 = pr.tables(name + '.dw')
        a2 = L.activate(r2, t2[0], t2[1], False)
        # tail: pw2 (+ shortcut) + shuffle
        out = pr.t(name + '.out')
        cat = L.unshuffle(out)                                        # raw values in concat order [shortcut | branch]
        y_ref = L.pw_forward(a2, pr.p(name + '.pw2.w'), pr.p(name + '.pw2.b'))
        rep['fwd'][name + '.pw2'] = rel(cat[..., sc:], y_ref)
        to = tuple(L.unshuffle(x) for x in pr.tables(name + '.out'))
        # (tables are exported in the shuffled order: compare the branch channels after un-shuffling)
        for nm, a, b in zip(('scale', 'shift', 'mean', 'inv'), (x[:, sc:] for x in to), L.bn_tables(cat[..., sc:], pr.p(name + '.pw2.g'), pr.p(name + '.pw2.be'))):
            rep['tables'][f'{name}.pw2:{nm}'] = ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
        if stride == 2:
            rs = pr.t(name + '.scdw')
            rep['fwd'][name + '.scdw'] = rel(rs, L.dw_forward(a_x, pr.p(name + '.scdw.w'), pr.p(name + '.scdw.b'), 2))
            _check_tables(rep, pr, name + '.scdw', rs, name + '.scdw')
            ts = pr.tables(name + '.scdw')
            a_s = L.activate(rs, ts[0], ts[1], False)
            rep['fwd'][name + '.scpw'] = rel(cat[..., :sc], L.pw_forward(a_s, pr.p(name + '.scpw.w'), pr.p(name + '.scpw.b')))
        else:
            # pass-through half: bit-exact copy of the raw input values and of their tables (channel_shuffle / split / concat
            # are pure index maps, core/architectures.py:109-118,122-123,143-144)
            rep['exact'][name + '.passthrough'] = float((cat[..., :sc] != x_raw[..., :sc]).sum().item())
            rep['exact'][name + '.passthrough_tables'] = float(sum((a[:, :sc] != b[:, :sc]).sum().item() for a, b in zip(to, x_tabs)))

        # ---------------------------------------------------------- backward of this unit (inputs: the stored d out)
        d_out = L.unshuffle(pr.t('grad:' + name + '.out'))            # d loss / d activated out, concat order
        S1o, S2o = (L.unshuffle(x) for x in pr.bsum(name + '.out'))
        dR_o, s1, s2 = L.bn_backward(d_out, cat, to[0], to[1], to[2], to[3], True, S1o, S2o)
        s1o_, s2o_ = s1, s2
        br = slice(sc, None)
        rep['sums'][name + '.pw2'] = max(rel(S1o[:, br], s1[:, br]), rel(S2o[:, br], s2[:, br]))
        d_a2, dW2 = L.pw_backward(dR_o[..., br], a2, pr.p(name + '.pw2.w'))
        rep['grad'][name + '.dw'] = rel(pr.t('grad:' + name + '.dw'), L.bf16r(d_a2))
        rep['pgrad'][name + '.pw2.w'] = rel(pr.g[name + '.pw2.w'], dW2)
        rep['pgrad'][name + '.pw2.g'] = rel(pr.g[name + '.pw2.g'], s2o_[:, br].sum(0))
        rep['pgrad'][name + '.pw2.be'] = rel(pr.g[name + '.pw2.be'], s1o_[:, br].sum(0))
        # depthwise
        S1, S2 = pr.bsum(name + '.dw')
        dR2, s1, s2 = L.bn_backward(pr.t('grad:' + name + '.dw'), r2, t2[0], t2[1], t2[2], t2[3], False, S1, S2)
        rep['sums'][name + '.dw'] = max(rel(S1, s1), rel(S2, s2))
        d_a1, dWd = L.dw_backward(dR2, a1, pr.p(name + '.dw.w'), stride)
        rep['grad'][name + '.pw1'] = rel(pr.t('grad:' + name + '.pw1'), L.bf16r(d_a1))
        rep['pgrad'][name + '.dw.w'] = rel(pr.g[name + '.dw.w'], dWd)
        rep['pgrad'][name + '.dw.g'] = rel(pr.g[name + '.dw.g'], s2.sum(0))
        rep['pgrad'][name + '.dw.be'] = rel(pr.g[name + '.dw.be'], s1.sum(0))
        # pw1
        S1, S2 = pr.bsum(name + '.pw1')
        dR1, s1, s2 = L.bn_backward(pr.t('grad:' + name + '.pw1'), r1, t1[0], t1[1], t1[2], t1[3], True, S1, S2)
        rep['sums'][name + '.pw1'] = max(rel(S1, s1), rel(S2, s2))
        d_xin, dW1 = L.pw_backward(dR1, xin, pr.p(name + '.pw1.w'))
        rep['pgrad'][name + '.pw1.w'] = rel(pr.g[name + '.pw1.w'], dW1)
        rep['pgrad'][name + '.pw1.g'] = rel(pr.g[name + '.pw1.g'], s2.sum(0))
        rep['pgrad'][name + '.pw1.be'] = rel(pr.g[name + '.pw1.be'], s1.sum(0))
        d_x = pr.t('grad:' + x_name)
        if stride == 2:
            S1, S2 = pr.bsum(name + '.scdw')
            dRs_in = dR_o[..., :sc]
            d_as, dWs = L.pw_backward(dRs_in, a_s, pr.p(name + '.scpw.w'))
            rep['grad'][name + '.scdw'] = rel(pr.t('grad:' + name + '.scdw'), L.bf16r(d_as))
            rep['pgrad'][name + '.scpw.w'] = rel(pr.g[name + '.scpw.w'], dWs)
            rep['pgrad'][name + '.scpw.g'] = rel(pr.g[name + '.scpw.g'], s2o_[:, :sc].sum(0))
            rep['pgrad'][name + '.scpw.be'] = rel(pr.g[name + '.scpw.be'], s1o_[:, :sc].sum(0))
            dRs, s1, s2 = L.bn_backward(pr.t('grad:' + name + '.scdw'), rs, ts[0], ts[1], ts[2], ts[3], False, S1, S2)
            rep['sums'][name + '.scdw'] = max(rel(S1, s1), rel(S2, s2))
            d_xs, dWsd = L.dw_backward(dRs, a_x, pr.p(name + '.scdw.w'), 2)
            rep['pgrad'][name + '.scdw.w'] = rel(pr.g[name + '.scdw.w'], dWsd)
            rep['pgrad'][name + '.scdw.g'] = rel(pr.g[name + '.scdw.g'], s2.sum(0))
            rep['pgrad'][name + '.scdw.be'] = rel(pr.g[name + '.scdw.be'], s1.sum(0))
            # two consumers (shortcut depthwise, then pw1): the sum is what the input tensor's gradient holds
            rep['grad'][x_name + '<-' + name] = rel(d_x, L.bf16r(L.bf16r(d_xs) + L.bf16r(d_xin)))
            rep['sums'][name + '.scpw'] = max(rel(S1o[:, :sc], s1o_[:, :sc]), rel(S2o[:, :sc], s2o_[:, :sc]))
        else:
            rep['grad'][x_name + '<-' + name] = rel(d_x[..., Cx // 2:], L.bf16r(d_xin))
            rep['exact'][name + '.passthrough_grad'] = float((d_x[..., :Cx // 2] != d_out[..., :sc]).sum().item())
        x_raw, x_tabs, x_clamp, x_name = out, pr.tables(name + '.out'), True, name + '.out'

    # ------------------------------------------------------------------ head conv + global average pool (:170-172)
    a_x = L.activate(x_raw, x_tabs[0], x_tabs[1], True)
    hd = pr.t('tower.head')
    rep['fwd']['tower.head'] = rel(hd, L.pw_forward(a_x, pr.p('tower.head.w'), pr.p('tower.head.b')))
    _check_tables(rep, pr, 'tower.head', hd, 'tower.head')
    th = pr.tables('tower.head')
    xs = L.per_slice(hd)
    a_h = (xs * L._bc(th[0], xs) + L._bc(th[1], xs)).clamp(0.0, 6.0).reshape(hd.shape)
    gap_ref = a_h.mean(dim=(1, 2)).reshape(4, B, -1)
    rep['fwd']['tower.gap'] = rel(pr.t('tower.gap'), gap_ref)
    dgap = pr.t('d.tower.gap').reshape(4 * B, 1, 1, -1)
    d_h_ref = L.bf16r(dgap / (hd.shape[1] * hd.shape[2])).expand_as(hd)
    rep['grad']['tower.head'] = rel(pr.t('grad:tower.head'), d_h_ref)
    S1, S2 = pr.bsum('tower.head')
    dRh, s1, s2 = L.bn_backward(pr.t('grad:tower.head'), hd, th[0], th[1], th[2], th[3], True, S1, S2)
    rep['sums']['tower.head'] = max(rel(S1, s1), rel(S2, s2))
    d_xh, dWh = L.pw_backward(dRh, a_x, pr.p('tower.head.w'))
    rep['grad'][x_name + '<-tower.head'] = rel(pr.t('grad:' + x_name), L.bf16r(d_xh))
    rep['pgrad']['tower.head.w'] = rel(pr.g['tower.head.w'], dWh)
    rep['pgrad']['tower.head.g'] = rel(pr.g['tower.head.g'], s2.sum(0))
    rep['pgrad']['tower.head.be'] = rel(pr.g['tower.head.be'], s1.sum(0))

    # ------------------------------------------------------------------ stem backward (max pool + BN + ReLU6 + conv weights)
    d_pool = pr.t('grad:tower.pool')
    d_a = L.maxpool_backward(d_pool, pidx, padded_hw, a_stem.shape)
    dRs, s1, s2 = L.bn_backward(d_a, stem, s_sc, s_sh, s_mu, s_inv, True)
    # the kernel folds the BatchNorm backward into the weight gradient in fp64 (no bf16 dR is ever stored): undo the rounding
    xs = L.per_slice(stem); dd = L.per_slice(d_a)
    z = L.f32r(xs * L._bc(s_sc, xs) + L._bc(s_sh, xs))
    dz = torch.where((z > 0) & (z < 6), dd, torch.zeros_like(dd))
    n = xs.numel() // (4 * xs.shape[-1])
    xhat = (xs - L._bc(s_mu, xs)) * L._bc(s_inv, xs)
    dR_exact = (L._bc(s_sc, xs) * (dz - L._bc(s1, xs) / n - xhat * L._bc(s2, xs) / n)).reshape(stem.shape)
    wl = pr.p('tower.stem.w').clone().requires_grad_(True)
    y = model.stem_conv(frames / 255.0, wl, torch.zeros(24, dtype=torch.float64))
    gw, = torch.autograd.grad(y, [wl], dR_exact)
    rep['pgrad']['tower.stem.w'] = rel(pr.g['tower.stem.w'], gw)
    rep['pgrad']['tower.stem.g'] = rel(pr.g['tower.stem.g'], s2.sum(0))
    rep['pgrad']['tower.stem.be'] = rel(pr.g['tower.stem.be'], s1.sum(0))
    return rep


def _worst(d):
    k = max(d, key=lambda k: d[k])
    return k, d[k]


def _assert_report(rep, tag):
    lines = [f'== {tag}']
    for cat in ('fwd', 'tables', 'sums', 'grad', 'pgrad', 'exact'):
        k, v = _worst(rep[cat])
        vals = sorted(rep[cat].values())
        lines.append(f'{cat:7s} n={len(vals):3d} median={vals[len(vals) // 2]:.3e} worst={v:.3e} ({k})')
    text = '\n'.join(lines)
    print(text)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, 'layerwise_parity.txt'), 'a') as f:
        f.write(text + '\n')
        for cat in ('fwd', 'grad', 'pgrad', 'sums'):
            for k, v in sorted(rep[cat].items(), key=lambda kv: -kv[1])[:8]:
                f.write(f'    {cat} {k} {v:.3e}\n')
    assert _worst(rep['exact'])[1] == 0.0, _worst(rep['exact'])            # index ops / max pool: bit-exact
    assert _worst(rep['fwd'])[1] <= FWD_TOL, _worst(rep['fwd'])
    assert _worst(rep['tables'])[1] <= TAB_TOL, _worst(rep['tables'])
    assert _worst(rep['grad'])[1] <= GRAD_TOL, _worst(rep['grad'])
    assert _worst(rep['pgrad'])[1] <= PGRAD_TOL, _worst(rep['pgrad'])
    assert _worst(rep['sums'])[1] <= GRAD_TOL, _worst(rep['sums'])


@pytest.mark.parametrize('weights', ['trained', 'random'])
def test_every_bf16_layer_matches_the_oracle(built_libs, weights):
    B = 8
    dyn, pol, val = C.trained_params(torch.float64) if weights == 'trained' else C.fresh_params(torch.float64)
    eng = _engine(B)
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, H, W, seed=101)), _dev(C.synthetic_batch(B, seed=102))
    sc = C.policy_step_engine(eng, obs, bt).cpu()
    torch.cuda.synchronize()
    assert torch.isfinite(sc[:10]).all()
    rep = _layerwise(eng, dyn, obs)
    _assert_report(rep, f'B={B} {weights} weights, policy pass')
    # the heads and the loss run in fp32 on the bf16 tower's output: given the library's own x512 the oracle's loss and head
    # gradients agree to the fp32 tolerance (PPO loss rtol 1e-3 asked by the north star for the perf path: met with margin)
    x512 = eng.x512.double().cpu().requires_grad_(True)
    h = {k: (t.clone().requires_grad_(True) if not k.endswith(('.mm', '.mv')) else t) for k, t in pol.items()}
    c = lambda t: t.detach().cpu().double()
    out = model.policy_forward(h, x512, c(bt['actions']), True)
    loss, scal = ppo.policy_objective(out, c(bt['adv']), c(bt['logp_old']), c(bt['true_speed']), c(bt['true_sim']), 0.2, 1.0)
    loss.backward()
    assert abs(sc[0].item() - loss.item()) <= 1e-4 * abs(loss.item()) + 1e-6
    rows = C.grad_report(eng.pol, eng.g_pol, {k: t.grad for k, t in h.items() if t.requires_grad})
    assert max(r[2] for r in rows) < 2e-3, sorted(rows, key=lambda r: -r[2])[:3]
    assert C.rel_l2(eng.d_x512, x512.grad) < 2e-3


@pytest.mark.parametrize('band', [3, 2])
def test_banded_depthwise_matches_the_oracle(built_libs, band):
    """The depthwise kernels work on row bands of a frame when it does not fit shared memory (180x240, BASELINE config 4):
    the test switch forces `band`-row bands on the 90x120 frames, where every layer is compared with the layer oracle at the
    same tolerances as the whole-frame kernels (halo rows re-read across bands, out-of-frame tile rows re-zeroed, each
    input row's gradient and each output row's weight-gradient share counted exactly once)."""
    from cdra import _lib
    lib = _lib.load()
    B = 8
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B)
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, H, W, seed=101)), _dev(C.synthetic_batch(B, seed=102))
    assert lib.cdra_debug_set(b'dw_band', band) == 0
    try:
        sc = C.policy_step_engine(eng, obs, bt).cpu()
        torch.cuda.synchronize()
        assert torch.isfinite(sc[:10]).all()
        rep = _layerwise(eng, dyn, obs)
    finally:
        lib.cdra_debug_set(b'dw_band', 0)
    _assert_report(rep, f'B={B} trained weights, depthwise bands of {band} rows')


def test_high_res_geometry_matches_the_oracle(built_libs):
    """BASELINE config 4 geometry (180x240): the stage-1 frames do not fit shared memory, the depthwise kernels band them on
    their own (no test switch), the stem / pool kernels run more bands; every layer against the layer oracle at the same
    tolerances as at 90x120."""
    B, h, w = 2, 180, 240
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B, h, w)
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, h, w, seed=121)), _dev(C.synthetic_batch(B, seed=122))
    sc = C.policy_step_engine(eng, obs, bt).cpu()
    torch.cuda.synchronize()
    assert torch.isfinite(sc[:10]).all()
    rep = _layerwise(eng, dyn, obs)
    _assert_report(rep, f'B={B} trained weights, 180x240 frames')


@pytest.mark.parametrize('route', ['fused=0', 'pwg=0', 'tc=0', 'fwd_tc=0'])
def test_every_kernel_route_matches_the_oracle(built_libs, route):
    """The library keeps alternative kernels for every pointwise layer (test switches of `cdra_debug_set`): `fused=0` = separate
    data-gradient (mma.sync) + tcgen05 weight-gradient kernels with the dR hand-off instead of the fused backward, `pwg=0` = the
    stage-3 / head layers on those kernels instead of the tcgen05 GEMM family, `tc=0` = mma.sync everywhere, `fwd_tc=0` =
    mma.sync forward.  Each route is held to the SAME per-layer oracle tolerances as the default one (this replaces the
    round-1 chained kernel-vs-kernel comparison and its 0.15 / 0.5 bounds)."""
    from cdra import _lib
    lib = _lib.load()
    key, val_ = route.split('=')
    B = 8
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B)
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, H, W, seed=101)), _dev(C.synthetic_batch(B, seed=102))
    assert lib.cdra_debug_set(key.encode(), int(val_)) == 0
    try:
        sc = C.policy_step_engine(eng, obs, bt).cpu()
        torch.cuda.synchronize()
        assert torch.isfinite(sc[:10]).all()
        rep = _layerwise(eng, dyn, obs)
    finally:
        lib.cdra_debug_set(key.encode(), -1)
    _assert_report(rep, f'B={B} trained weights, kernel route {route}')


def test_full_size_layer_slice_bf16(built_libs):
    """BASELINE config-2 minibatch (B = 512): the same per-layer comparison on one slice of rows of one stage-1, one
    stage-2 and one stage-3 pointwise layer (forward output, stored input gradient) -- the grids, tile schedules and TMA ring
    depths differ from the B = 8 case."""
    B = 512
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B)
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, H, W, seed=111)), _dev(C.synthetic_batch(B, seed=112))
    C.policy_step_engine(eng, obs, bt)
    torch.cuda.synchronize()
    p = lambda n: dyn[n].double()
    for name in ('tower.s1.u2', 'tower.s2.u5', 'tower.s3.u1'):
        cx = {'tower.s1.u2': 'tower.s1.u1.out', 'tower.s2.u5': 'tower.s2.u4.out', 'tower.s3.u1': 'tower.s3.u0.out'}[name]
        x = eng.tensor(cx)                                             # [4B,h,w,C] fp32 on the device
        Cx = x.shape[-1]
        aff = eng.tensor('aff:' + cx).squeeze(-1).double().cpu()
        a1t = eng.tensor('aff:' + name + '.pw1').squeeze(-1).double().cpu()
        b1t = eng.tensor('bnp:' + name + '.pw1').squeeze(-1).double().cpu()
        bs = eng.tensor('bsum:' + name + '.pw1').squeeze(-1).double().cpu()
        r1 = eng.tensor(name + '.pw1'); g1 = eng.tensor('grad:' + name + '.pw1'); gx = eng.tensor('grad:' + cx)
        n = r1.shape[0] // 4 * r1.shape[1] * r1.shape[2]
        for t in (0, 3):
            for b in (0, B // 2 + 3, B - 1):                         # frames at the start / middle / end of a slice's tile schedule
                f = t * B + b
                xf = x[f].double().cpu()
                z = L.bf16r(L.f32r(xf * aff[t, :, 0] + aff[t, :, 1])).clamp(0, 6)
                xin = z[..., Cx // 2:]
                ref = L.pw_forward(xin, p(name + '.pw1.w'), p(name + '.pw1.b'))
                assert C.rel_l2(r1[f], ref) <= FWD_TOL, (name, f)
                raw = r1[f].double().cpu(); dA = g1[f].double().cpu()
                zz = L.f32r(raw * a1t[t, :, 0] + a1t[t, :, 1])
                dz = torch.where((zz > 0) & (zz < 6), dA, torch.zeros_like(dA))
                xhat = (raw - b1t[t, :, 0]) * b1t[t, :, 1]
                dR = L.bf16r(a1t[t, :, 0] * (dz - bs[t, :, 0] / n - xhat * bs[t, :, 1] / n))
                d_in, _ = L.pw_backward(dR, xin, p(name + '.pw1.w'))
                assert C.rel_l2(gx[f][..., Cx // 2:], L.bf16r(d_in)) <= GRAD_TOL, (name, f)
        # the unit tail (pw2 + channel shuffle + pass-through) of the same unit: branch channels against the layer oracle,
        # pass-through channels bit-exact
        r2 = eng.tensor(name + '.dw'); out = eng.tensor(name + '.out')
        a2t = eng.tensor('aff:' + name + '.dw').squeeze(-1).double().cpu()
        for t, b in ((0, 1), (2, B - 2)):
            f = t * B + b
            a2 = L.bf16r(L.f32r(r2[f].double().cpu() * a2t[t, :, 0] + a2t[t, :, 1]))
            cat = L.unshuffle(out[f].double().cpu())
            assert C.rel_l2(cat[..., Cx // 2:], L.pw_forward(a2, p(name + '.pw2.w'), p(name + '.pw2.b'))) <= FWD_TOL, (name, f)
            assert torch.equal(cat[..., :Cx // 2], x[f].double().cpu()[..., :Cx // 2]), (name, f)
