"""CPU logic checks of the library's plain-CUDA kernels through the SIMT emulator (tests/emu/), against
the oracle, at sizes the emulator finishes in seconds.  These exercise indexing / tiling / channel-map /
BatchNorm bookkeeping logic only; the parity tests proper run on the GPU (test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

from oracle import model, ppo, spec
from tests import common as C

B, H, W = 3, 42, 58      # stem 20x28 -> pool 10x14 -> 5x7 -> 3x4 -> 2x2


@pytest.fixture(scope='module')
def eng(built_libs):
    from tests.emu.engine import EmuEngine
    e = EmuEngine(B, H, W, dtype='f32', image_u8=True, device='cpu')
    return e


@pytest.fixture(scope='module')
def params():
    return C.fresh_params(torch.float64)


def test_arena_layout_matches_oracle_spec(eng):
    for arena, state, pspec in ((eng.dyn, eng.dyn_state, spec.dynamics_params()), (eng.pol, eng.pol_state, spec.head_params('policy')),
                                (eng.val, eng.val_state, spec.head_params('value'))):
        tr, nt, st, ns = spec.split_layout(pspec)
        assert arena.names == list(tr) and arena.size == nt
        assert state.names == list(st) and state.size == ns
        for n, (off, shape) in tr.items():
            i = arena.index[n]
            assert arena.offsets[i] == off and arena.shapes[i] == tuple(shape)


def test_forward_taps_and_moving_stats(eng, params):
    dyn, pol, val = params
    C.load_engine(eng, dyn, pol, val)
    obs = C.synthetic_obs(B, H, W, seed=5)
    out = eng.dynamics_forward(obs).clone()
    taps, bs = {}, model.BNState()
    ref = model.dynamics_forward(dyn, C.oracle_obs(obs), True, bs, taps)
    for k in ('tower.stem', 'tower.pool', 'tower.s1.u0.pw1', 'tower.s1.u0.dw', 'tower.s1.u0.scdw', 'tower.s1.u1.pw1',
              'tower.s2.u0.pw1', 'tower.s3.u3.dw', 'tower.head'):
        assert C.rel_max(eng.tensor(k)[:B], taps[k]) < 2e-4, k
    assert C.rel_max(eng.tensor('tower.gap'), taps['tower.gap']) < 5e-4
    assert C.rel_max(eng.tensor('dynamics_in'), taps['dynamics_in']) < 5e-4
    assert C.rel_max(out, ref) < 1e-3
    new = bs.apply_moving(dyn)
    got = eng.dyn_state.to_dict()
    assert max((got[k].double() - new[k]).abs().max().item() for k in got) < 1e-5


def test_inference_mode_uses_moving_statistics(eng, params):
    """CARLANetwork.dynamics_predict (training=False, core/networks.py:206-208)"""
    dyn, pol, val = params
    C.load_engine(eng, dyn, pol, val)
    obs = C.synthetic_obs(B, H, W, seed=15)
    state_before = eng.dyn_state.flat.clone()
    out = eng.dynamics_forward(obs, training=False).clone()
    ref = model.dynamics_forward(dyn, C.oracle_obs(obs), training=False)
    assert C.rel_max(out, ref) < 1e-4
    assert torch.equal(eng.dyn_state.flat, state_before)           # inference never touches the moving averages
    z = torch.zeros(B, 2)
    eng.value_head(out, z, z[:, :1].contiguous(), z[:, :1].contiguous(), training=False, backward=False)
    v = model.value_forward(val, ref, training=False)
    assert C.rel_max(eng.head_out.view(-1)[:B * 4].view(B, 4)[:, :2], v['value']) < 1e-4
    eng.policy_head(out, torch.full((B, 2), 0.5), z, z[:, 0].contiguous(), z[:, :1].contiguous(), z[:, :1].contiguous(),
                    training=False, backward=False)
    pi = model.policy_forward(pol, ref, torch.full((B, 2), 0.5, dtype=torch.float64), training=False)
    ho = eng.head_out.view(B, 8)
    assert C.rel_max(ho[:, 0:2], pi['alpha']) < 1e-4 and C.rel_max(ho[:, 2:4], pi['beta']) < 1e-4
    assert C.rel_max(ho[:, 6:8], pi['log_prob']) < 1e-4


def test_channel_shuffle_is_bit_exact(eng, params):
    """the shortcut half of a stride-1 unit is a pure index permutation of the input's left half"""
    dyn, pol, val = params
    C.load_engine(eng, dyn, pol, val)
    eng.dynamics_forward(C.synthetic_obs(B, H, W, seed=6))
    x, y = eng.tensor('tower.s1.u0.out'), eng.tensor('tower.s1.u1.out')
    c = x.shape[-1]
    perm = model.shuffle_perm(c)                       # out[j] = concat[perm[j]]
    for j, q in enumerate(perm):
        if q < c // 2:                                  # concat position q < C/2 is the shortcut = x[..., q]
            assert torch.equal(y[..., j], x[..., q])


def test_policy_pass_gradients(eng, params):
    dyn, pol, val = params
    C.load_engine(eng, dyn, pol, val)
    obs, bt = C.synthetic_obs(B, H, W, seed=7), C.synthetic_batch(B, seed=8)
    sc = C.policy_step_engine(eng, obs, bt)
    ref = C.policy_step_oracle(dyn, pol, obs, bt)
    assert abs(sc[0].item() - ref['loss'].item()) < 1e-4 * max(1.0, abs(ref['loss'].item()))
    names = ['loss_total', 'loss_policy', 'loss_entropy', 'loss_speed_policy', 'loss_similarity_policy', 'ratio', 'log_prob',
             'entropy', 'speed_pi', 'similarity_pi']
    for i, n in enumerate(names):
        assert abs(sc[i].item() - ref['scalars'][n].item()) < 2e-4 * max(1.0, abs(ref['scalars'][n].item())), n
    rows = C.grad_report(eng.pol, eng.g_pol, ref['g_head'])
    assert max(r[2] for r in rows) < 1e-2      # BatchNorm over B = 3 rows amplifies fp32 rounding
    rows = C.grad_report(eng.dyn, eng.g_dyn, ref['g_dyn'])
    tail = [r for r in rows if not r[0].startswith('tower.')]
    assert max(r[2] for r in tail) < 2e-2, sorted(tail, key=lambda r: -r[2])[:3]
    # ReLU6 masks can flip between fp32 kernels and the fp64 oracle at this tiny batch (BatchNorm over
    # <= 12 values amplifies rounding): judge the tower by relative L2, which a single flip barely moves
    l2 = sorted(r[1] for r in rows)
    assert l2[len(l2) // 2] < 2e-2 and l2[-1] < 0.25, (l2[len(l2) // 2], l2[-1])


def test_value_pass_gradients(eng, params):
    dyn, pol, val = params
    C.load_engine(eng, dyn, pol, val)
    obs, bt = C.synthetic_obs(B, H, W, seed=9), C.synthetic_batch(B, seed=10)
    sc = C.value_step_engine(eng, obs, bt)
    ref = C.value_step_oracle(dyn, val, obs, bt)
    assert abs(sc[0].item() - ref['loss'].item()) < 1e-4 * max(1.0, abs(ref['loss'].item()))
    rows = C.grad_report(eng.val, eng.g_val, ref['g_head'])
    assert max(r[2] for r in rows) < 1e-2
    rows = C.grad_report(eng.dyn, eng.g_dyn, ref['g_dyn'])
    l2 = sorted(r[1] for r in rows)
    assert l2[len(l2) // 2] < 2e-2 and l2[-1] < 0.25


def test_gae_returns_exponents_bit_exact(eng):
    rng = np.random.RandomState(0)
    bs, T = 6, 41
    rew = (rng.randn(bs, T) * 2 + 1).clip(-10, 30).astype('f')
    vbe = np.stack([rng.rand(bs, T) * 2 - 1, rng.rand(bs, T) * 6], -1).astype('f')
    last = np.stack([rng.rand(bs) * 2 - 1, rng.rand(bs) * 6], -1).astype('f')
    last[0] = 0                                         # terminal state: CARLANetwork.last_value zeros (networks.py:171)
    rb, adv = eng.gae(torch.tensor(rew), torch.tensor(vbe), torch.tensor(last), 0.9999, 0.999, 2.0)
    for i in range(bs):
        r_ref, a_ref, _ = ppo.end_trajectory(rew[i], vbe[i], last[i], 0.9999, 0.999, 2.0)
        assert np.array_equal(rb[i].numpy(), r_ref)     # base and exponent bit-exact (sequential fp64 filter)
        assert np.abs(adv[i].numpy() - a_ref).max() < 5e-7 * 2.0


def test_clip_adam_three_steps(eng):
    arena = eng.val
    g = torch.Generator().manual_seed(3)
    p0 = {n: torch.randn(s, generator=g) for n, s in zip(arena.names, arena.shapes)}
    g0 = {n: torch.randn(s, generator=g) * (3.0 if i % 2 else 0.01) for i, (n, s) in enumerate(zip(arena.names, arena.shapes))}
    arena.load_dict(p0)
    for n in arena.names:
        arena.view(n, eng.g_val).copy_(g0[n])
    eng.adam['val'][0].zero_(); eng.adam['val'][1].zero_(); eng.adam_step['val'] = 0
    pr = {k: v.clone().double() for k, v in p0.items()}
    m = {k: torch.zeros_like(v) for k, v in pr.items()}
    v = {k: torch.zeros_like(t) for k, t in pr.items()}
    for step in (1, 2, 3):
        eng.clip_adam('val', 3e-4, clip_norm=1.0)
        ppo.apply_step(pr, {k: t.double() for k, t in g0.items()}, m, v, step, 3e-4, clip=1.0)
    got = arena.to_dict()
    assert max((got[k].double() - pr[k]).abs().max().item() for k in got) < 2e-6


def test_gather_rows(eng):
    src = torch.arange(7 * 48, dtype=torch.uint8).view(7, 48).contiguous()
    idx = torch.tensor([5, 0, 3, 3], dtype=torch.int64)
    out = torch.empty(4, 48, dtype=torch.uint8)
    eng.gather_rows(src, idx, out)
    assert torch.equal(out, src[idx])
    src = torch.randn(9, 5)
    out = torch.empty(3, 5)
    eng.gather_rows(src, torch.tensor([8, 1, 4]), out)
    assert torch.equal(out, src[[8, 1, 4]])
    # every tensor of a minibatch in one launch (ragged row sizes: 16-byte vector path, byte path, 4-byte rows)
    srcs = [torch.arange(9 * 4096, dtype=torch.uint8).view(9, 4096).contiguous(), torch.randn(9, 5), torch.randn(9, 1), torch.randn(9, 4, 9)]
    idx = torch.tensor([8, 1, 4, 4, 0], dtype=torch.int64)
    outs = [torch.empty((5,) + tuple(t.shape[1:]), dtype=t.dtype) for t in srcs]
    eng.gather_rows_multi(srcs, idx, outs)
    for t, o in zip(srcs, outs):
        assert torch.equal(o, t[idx])


def test_policy_head_reparameterized_sample_gradient(eng, params):
    """PolicyNetwork.call evaluates log pi at a REPARAMETERISED sample of the new policy (core/networks.py:97-100; TFP Beta
    is FULLY_REPARAMETERIZED [lib]): with the sample's pathwise derivatives handed over (`actions_jac`), the fused head kernel
    must reproduce the oracle's gradient through the sample; and the jacobian itself is checked against the defining property
    of a reparameterised sample, d/dalpha E[f(x)] = E[f'(x) dx/dalpha]."""
    dyn, pol, val = params
    C.load_engine(eng, dyn, pol, val)
    g = torch.Generator().manual_seed(11)
    x512 = torch.randn(B, 512, generator=g)
    bt = C.synthetic_batch(B, seed=12)
    alpha = 1.01 + 3.0 * torch.rand(B, 2, generator=g, dtype=torch.float64)
    beta = 1.01 + 3.0 * torch.rand(B, 2, generator=g, dtype=torch.float64)
    act, jac = model.beta_sample_reparameterized(alpha, beta, generator=g)
    act[0, 0] = 1.0                       # a clipped sample: tf.clip_by_value stops its gradient
    bt['actions'], bt['actions_jac'] = act.float().contiguous(), jac.float().contiguous()
    grads = {}
    for use_jac in (True, False):
        sc = eng.policy_head(x512, bt['actions'], bt['logp_old'], bt['adv'], bt['true_speed'], bt['true_sim'], 0.2, 1.0,
                             actions_jac=bt['actions_jac'] if use_jac else None).clone()
        h = {k: (t.clone().requires_grad_(True) if not k.endswith(('.mm', '.mv')) else t) for k, t in pol.items()}
        xl = x512.double().requires_grad_(True)
        out = model.policy_forward(h, xl, bt['actions'].double(), True, None, bt['actions_jac'].double() if use_jac else None)
        loss, _ = ppo.policy_objective(out, bt['adv'].double(), bt['logp_old'].double(), bt['true_speed'].double(), bt['true_sim'].double(), 0.2, 1.0)
        loss.backward()
        assert abs(sc[0].item() - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
        rows = C.grad_report(eng.pol, eng.g_pol, {k: t.grad for k, t in h.items() if t.requires_grad})
        assert max(r[2] for r in rows) < 1e-2, sorted(rows, key=lambda r: -r[2])[:3]
        assert C.rel_max(eng.d_x512, xl.grad) < 1e-2
        grads[use_jac] = eng.g_pol.clone()
    assert C.rel_l2(grads[True], grads[False]) > 1e-2          # the pathwise term is not negligible
    # the jacobian is a pathwise derivative: compare E[x * dx/dalpha-weighted] estimates with the closed form of the mean,
    # d/dalpha E[x] = beta / (alpha + beta)^2, d/dbeta E[x] = -alpha / (alpha + beta)^2
    a0, b0 = torch.full((200000,), 2.5, dtype=torch.float64), torch.full((200000,), 1.7, dtype=torch.float64)
    x, j = model.beta_sample_reparameterized(a0, b0, generator=g)
    assert abs(j[:, 0].mean().item() - 1.7 / (4.2 ** 2)) < 2e-3 and abs(j[:, 1].mean().item() + 2.5 / (4.2 ** 2)) < 2e-3
