"""Shared helpers for the parity tests: seeded synthetic inputs, oracle <-> engine glue,
error metrics."""
import os

import numpy as np
import torch

from oracle import model, ppo, spec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def synthetic_obs(B, H, W, seed=1234, device='cpu', u8=True):
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (B, 4, H, W, 3), dtype=torch.uint8, generator=g)
    road = torch.cat([(torch.rand(B, 4, 3, generator=g) < 0.2).float(), 0.3 + 0.6 * torch.rand(B, 4, 1, generator=g),
                      torch.nn.functional.one_hot(torch.randint(0, 5, (B, 4), generator=g), 5).float()], dim=-1)
    veh = torch.cat([torch.rand(B, 4, 1, generator=g) * 2 - 1, torch.rand(B, 4, 3, generator=g)], dim=-1)
    nav = torch.sort(torch.rand(B, 4, 5, generator=g) * 25, dim=-1).values
    obs = dict(state_image=img if u8 else img.float() / 255, state_road=road, state_vehicle=veh, state_navigation=nav)
    return {k: v.contiguous().to(device) for k, v in obs.items()}


def synthetic_batch(B, seed=99, device='cpu'):
    g = torch.Generator().manual_seed(seed)
    d = dict(actions=torch.distributions.Beta(2.0, 2.0).sample((B, 2)).clamp(1e-6, 1 - 1e-6) if False else
             torch.rand(B, 2, generator=g).clamp(1e-4, 1 - 1e-4),
             adv=torch.randn(B, generator=g), logp_old=0.5 * torch.randn(B, 2, generator=g),
             true_speed=0.3 * torch.rand(B, 1, generator=g), true_sim=torch.rand(B, 1, generator=g) * 2 - 1,
             returns=torch.stack([torch.rand(B, generator=g) * 2 - 1, torch.randint(0, 5, (B,), generator=g).float()], dim=1))
    return {k: v.contiguous().to(device) for k, v in d.items()}


def oracle_obs(obs, dtype=torch.float64):
    o = {k: v.detach().cpu().to(dtype) for k, v in obs.items()}
    if obs['state_image'].dtype == torch.uint8:
        o['state_image'] = obs['state_image'].detach().cpu().to(dtype) / 255.0
    return o


def fresh_params(dtype=torch.float64, seed=1):
    dyn = model.randomize_bn(model.init_params(spec.dynamics_params(), seed=seed, dtype=dtype), seed=seed + 10)
    pol = model.randomize_bn(model.init_params(spec.head_params('policy'), seed=seed + 1, dtype=dtype), seed=seed + 11)
    val = model.randomize_bn(model.init_params(spec.head_params('value'), seed=seed + 2, dtype=dtype), seed=seed + 12)
    return dyn, pol, val


def trained_params(dtype=torch.float64):
    """The reference's shipped stage-s5-curriculum agent (tests/golden/ckpt_s5_curriculum.npz, fp16-rounded)."""
    z = np.load(os.path.join(GOLDEN, 'ckpt_s5_curriculum.npz'))
    out = []
    for prefix in ('dyn/', 'pol/', 'val/'):
        out.append({k[len(prefix):]: torch.from_numpy(z[k].astype(np.float32)).to(dtype) for k in z.files if k.startswith(prefix)})
    return tuple(out)


def load_engine(eng, dyn, pol, val):
    eng.dyn.load_dict(dyn); eng.dyn_state.load_dict(dyn)
    eng.pol.load_dict(pol); eng.pol_state.load_dict(pol)
    eng.val.load_dict(val); eng.val_state.load_dict(val)


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def policy_step_engine(eng, obs, bt, clip=0.2, ent=1.0):
    x = eng.dynamics_forward(obs)
    sc = eng.policy_head(x, bt['actions'], bt['logp_old'], bt['adv'], bt['true_speed'], bt['true_sim'], clip, ent,
                         actions_jac=bt.get('actions_jac')).clone()
    eng.dynamics_backward(obs, eng.d_x512)
    return sc


def value_step_engine(eng, obs, bt):
    x = eng.dynamics_forward(obs)
    sc = eng.value_head(x, bt['returns'], bt['true_speed'], bt['true_sim']).clone()
    eng.dynamics_backward(obs, eng.d_x512)
    return sc


def policy_step_oracle(dyn, pol, obs, bt, clip=0.2, ent=1.0, dtype=torch.float64):
    c = lambda t: t.detach().cpu().to(dtype)
    return ppo.policy_pass(dyn, pol, oracle_obs(obs, dtype), c(bt['actions']), c(bt['adv']), c(bt['logp_old']),
                           c(bt['true_speed']), c(bt['true_sim']), clip, ent,
                           actions_jac=c(bt['actions_jac']) if 'actions_jac' in bt else None)


def value_step_oracle(dyn, val, obs, bt, dtype=torch.float64):
    c = lambda t: t.detach().cpu().to(dtype)
    return ppo.value_pass(dyn, val, oracle_obs(obs, dtype), c(bt['returns']), c(bt['true_speed']), c(bt['true_sim']))


# tensors whose true gradient is analytically zero (a per-channel constant added right before a
# training-mode BatchNorm) -- the reference computes rounding noise for them.
def zero_grad_tensor(name, ref_grad):
    return ref_grad.abs().max().item() < 1e-9


def grad_report(eng_arena, eng_flat, ref_grads):
    """-> list of (name, rel_l2, rel_max, ref_max) for tensors with a non-trivial gradient."""
    mine = eng_arena.to_dict(eng_flat)
    rows = []
    for k, g in ref_grads.items():
        if zero_grad_tensor(k, g):
            continue
        rows.append((k, rel_l2(mine[k], g), rel_max(mine[k], g), g.abs().max().item()))
    return rows
