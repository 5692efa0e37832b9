"""Oracle restatement of the PPO numerics (losses, GAE/returns, clipping, Adam).  Test infrastructure.

GAE / returns call the *same* third-party routine the reference calls (`scipy.signal.lfilter`,
rl/utils.py:57-59); everything else is restated from the cited lines.
"""
import numpy as np
import scipy.signal
import torch

from . import model, spec


# ----------------------------------------------------------------------------- losses
def policy_objective(out, advantages, old_log_prob, true_speed, true_similarity, clip_ratio, ent_coef):
    """core/carla_agent.py:394-428.  `out` = model.policy_forward(...).  Returns (total, scalars)."""
    log_prob = out['log_prob']
    entropy = out['entropy'].mean()
    entropy_penalty = ent_coef * entropy
    ratio = torch.exp(log_prob - old_log_prob).mean(dim=1)
    min_adv = torch.where(advantages > 0.0, (1.0 + clip_ratio) * advantages, (1.0 - clip_ratio) * advantages)
    speed_loss = 0.5 * ((true_speed - out['speed']) ** 2).mean(dim=-1).mean()
    sim_loss = 0.5 * ((true_similarity - out['similarity']) ** 2).mean(dim=-1).mean()
    policy_loss = -torch.minimum(ratio * advantages, min_adv).mean()
    total = policy_loss - entropy_penalty + speed_loss + sim_loss
    scalars = dict(loss_total=total, loss_policy=policy_loss, loss_entropy=entropy_penalty,
                   loss_speed_policy=speed_loss, loss_similarity_policy=sim_loss, ratio=ratio.mean(),
                   log_prob=log_prob.mean(), entropy=entropy, speed_pi=out['speed'].mean(),
                   similarity_pi=out['similarity'].mean())
    return total, scalars


def value_objective(out, returns, true_speed, true_similarity, exp_scale=6.0):
    """core/carla_agent.py:469-486."""
    values = out['value']
    base_loss = ((returns[:, 0] - values[:, 0]) ** 2).mean()
    exp_loss = ((returns[:, 1] - values[:, 1]) ** 2).mean()
    value_loss = 0.25 * base_loss + exp_loss / (exp_scale ** 2)
    speed_loss = ((true_speed - out['speed']) ** 2).mean(dim=-1).mean()
    sim_loss = ((true_similarity - out['similarity']) ** 2).mean(dim=-1).mean()
    total = (value_loss + speed_loss + sim_loss) * 0.25
    scalars = dict(loss_total=total, loss_v=value_loss, loss_speed_value=speed_loss,
                   loss_similarity_value=sim_loss, speed_v=out['speed'].mean(),
                   similarity_v=out['similarity'].mean())
    return total, scalars


# ----------------------------------------------------------------------------- returns / GAE
def discount_cumsum(x, discount):
    """rl/utils.py:57-59 verbatim semantics (scipy promotes to float64)."""
    return scipy.signal.lfilter([1.0], [1.0, float(-discount)], x[::-1], axis=0)[::-1]


def decompose_number(num):
    """rl/utils.py:140-151 on a float32 scalar (the reference runs it through tf.map_fn on fp32)."""
    num = np.float32(num)
    exponent = 0
    ten = np.float32(10.0)
    while abs(num) > np.float32(1.0):
        num = np.float32(num / ten)
        exponent += 1
    return num, np.float32(exponent)


def sp_norm(x, eps=1e-3):
    """rl/utils.py:344-349 in float32."""
    x = x.astype(np.float32)
    pos = x * (x > 0.0).astype(np.float32)
    neg = x * (x < 0.0).astype(np.float32)
    return pos / np.float32(x.max() + np.float32(eps)) + neg / np.float32(-(x.min() - np.float32(eps)))


def end_trajectory(rewards, values_be, last_value_be, gamma, lambda_, scale):
    """PPOMemory.end_trajectory + compute_returns + compute_advantages for ONE trajectory
    (rl/agents/ppo.py:692-727).  rewards [T] fp32, values_be [T,2] fp32, last_value_be [2].
    Returns (returns_be [T,2] fp32, advantages [T] fp32, raw returns [T] fp32)."""
    rewards = np.asarray(rewards, np.float32)
    values_be = np.asarray(values_be, np.float32)
    last = np.asarray(last_value_be, np.float32)

    def pow10(e):
        # tf.pow(10.0, e) on CPU is Eigen's scalar std::pow(float, float) [lib], i.e. glibc powf, which is
        # correctly rounded; numpy's float32 `power` is a SIMD kernel that is up to 1 ulp off, so the
        # oracle evaluates in float64 and rounds once.
        return np.power(10.0, np.asarray(e, np.float64)).astype(np.float32)

    v_T = np.float32(last[0] * pow10(last[1]))
    r = np.concatenate([rewards, [v_T]]).astype(np.float32)
    vbe = np.concatenate([values_be, last[None]], axis=0)
    returns = discount_cumsum(r, gamma)[:-1].astype(np.float32)
    dec = [decompose_number(x) for x in returns]
    returns_be = np.array(dec, np.float32).reshape(-1, 2)
    v = (vbe[:, 0] * pow10(vbe[:, 1])).astype(np.float32)
    g = np.float32(gamma)
    deltas = ((r[:-1] + g * v[1:]).astype(np.float32) - v[:-1]).astype(np.float32)
    adv = discount_cumsum(deltas, gamma * lambda_).astype(np.float32)
    adv_n = (sp_norm(adv) * np.float32(scale)).astype(np.float32)
    return returns_be, adv_n, returns


# ----------------------------------------------------------------------------- optimiser
def clip_by_norm(g, clip):
    """tf.clip_by_norm (rl/utils.py:120-121) [lib]: g*clip / max(||g||, clip)."""
    l2sum = (g * g).sum()
    l2 = torch.sqrt(torch.where(l2sum > 0, l2sum, torch.ones_like(l2sum)))
    return g * clip / torch.maximum(l2, torch.tensor(clip, dtype=g.dtype))


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras Adam._resource_apply_dense [lib]: eps is added to the uncorrected sqrt(v).
    `step` is the 1-based iteration count.  In place on p, m, v."""
    lr_t = lr * np.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    p.sub_(lr_t * m / (v.sqrt() + eps))


# ----------------------------------------------------------------------------- one SGD step of each pass
def _leaf(params):
    return {k: (t.clone().requires_grad_(True) if not k.endswith(('.mm', '.mv')) else t) for k, t in params.items()}


def policy_pass(dyn, pol, obs, actions_eval, advantages, old_log_prob, true_speed, true_sim,
                clip_ratio=0.2, ent_coef=1.0, actions_jac=None):
    """CARLAgent.get_policy_gradients (core/carla_agent.py:351-373): loss + grads wrt policy head and
    dynamics.  Returns dict(loss, scalars, g_dyn, g_pol, x512, bn_dyn, bn_pol)."""
    d, h = _leaf(dyn), _leaf(pol)
    bs_d, bs_h = model.BNState(), model.BNState()
    x512 = model.dynamics_forward(d, obs, True, bs_d)
    out = model.policy_forward(h, x512, actions_eval, True, bs_h, actions_jac)
    loss, scalars = policy_objective(out, advantages, old_log_prob, true_speed, true_sim, clip_ratio, ent_coef)
    loss.backward()
    return dict(loss=loss.detach(), scalars={k: v.detach() for k, v in scalars.items()},
                g_dyn={k: t.grad for k, t in d.items() if t.requires_grad},
                g_head={k: t.grad for k, t in h.items() if t.requires_grad},
                x512=x512.detach(), out={k: v.detach() for k, v in out.items()},
                bn_dyn=bs_d.apply_moving(dyn), bn_head=bs_h.apply_moving(pol))


def value_pass(dyn, val, obs, returns, true_speed, true_sim):
    """CARLAgent.get_value_gradients (core/carla_agent.py:430-452)."""
    d, h = _leaf(dyn), _leaf(val)
    bs_d, bs_h = model.BNState(), model.BNState()
    x512 = model.dynamics_forward(d, obs, True, bs_d)
    out = model.value_forward(h, x512, True, bs_h)
    loss, scalars = value_objective(out, returns, true_speed, true_sim)
    loss.backward()
    return dict(loss=loss.detach(), scalars={k: v.detach() for k, v in scalars.items()},
                g_dyn={k: t.grad for k, t in d.items() if t.requires_grad},
                g_head={k: t.grad for k, t in h.items() if t.requires_grad},
                x512=x512.detach(), out={k: v.detach() for k, v in out.items()},
                bn_dyn=bs_d.apply_moving(dyn), bn_head=bs_h.apply_moving(val))


def apply_step(params, grads, m, v, step, lr, clip=None):
    """apply_{policy,value}_gradients (rl/agents/ppo.py:238-275): per-tensor clip_by_norm then Adam;
    apply_dynamics_gradients (core/carla_agent.py:386-388): Adam, no clipping (clip=None)."""
    for k, g in grads.items():
        if clip is not None:
            g = clip_by_norm(g, clip)
        adam_step(params[k], g, m[k], v[k], step, lr)
