"""The reference-facing API on the GPU (no logic-check build anywhere): `CARLAgent.learn()` / `update()` on a synthetic
rollout (core/carla_agent.py:129-145, rl/agents/ppo.py:190-226,464-568), the rollout path `CARLANetwork.predict` /
`predict_last_value` (core/networks.py:178-221: eval-mode BatchNorm, Beta sampling, value = base * 10^exp) against the oracle
on the reference's trained stage-s5-curriculum agent loaded from a checkpoint in the REFERENCE'S on-disk format, and the
device-resident rollout memory feeding the update."""
import os

import numpy as np
import pytest
import torch

from oracle import ckpt, model, spec
from tests import common as C
from tests.golden import tf_bundle_writer as W

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason='needs a CUDA device')]

H, Wd = 90, 120


def _agent(tmp_path, batch_size, dtype='bf16', name='t', **kw):
    from core import CARLAgent, SyntheticCARLAEnvironment
    env = SyntheticCARLAEnvironment(image_shape=(H, Wd, 3), image_uint8=True, seed=1)
    base = dict(seed=7, skip_data=1, drop_batch_remainder=True, log_mode='summary', policy_lr=3e-4, value_lr=3e-4,
                dynamics_lr=3e-4, entropy_regularization=1.0, clip_ratio=0.2, gamma=0.9999, lambda_=0.999, aug_intensity=0.0)
    base.update(kw)
    return CARLAgent(env, batch_size=batch_size, name=name, weights_dir=str(tmp_path / 'w'), evaluation_dir=str(tmp_path / 'e'),
                     network=dict(dtype=dtype), **base)


def _reference_style_checkpoint(folder):
    dyn, pol, val = C.trained_params(torch.float32)
    os.makedirs(folder, exist_ok=True)
    for fname, params, order in (('dynamics_model', dyn, ckpt.dynamics_layer_order()), ('policy_net', pol, ckpt.head_layer_order('policy')),
                                 ('value_net', val, ckpt.head_layer_order('value'))):
        W.write_bundle(os.path.join(folder, fname), W.keras_checkpoint_tensors({k: v.numpy() for k, v in params.items()}, order))
    with open(os.path.join(folder, 'config.json'), 'w') as f:          # the shipped folders carry one (rl/agents/ppo.py:601-616)
        f.write('{"policy_lr": {"step": 3}, "value_lr": {"step": 3}, "dynamics_lr": {"step": 3}}')
    return dyn, pol, val


def test_learn_and_update_on_cuda(built_libs, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                              # TensorBoard event files go to ./logs like the reference's
    agent = _agent(tmp_path, batch_size=8)
    net = agent.network
    assert net.device.type == 'cuda' and net.dtype == 'bf16'
    p0, v0, d0 = net.policy.flat.clone(), net.value.flat.clone(), net.dynamics.flat.clone()
    agent.learn(episodes=1, timesteps=33, save_every='end', close=False)
    # 33 transitions, skip(1) -> 32 -> 4 minibatches of 8 per pass; the trunk steps once per pass
    assert net.engine.adam_step == dict(dyn=8, pol=4, val=4)
    for a, b in ((net.policy.flat, p0), (net.value.flat, v0), (net.dynamics.flat, d0)):
        assert torch.isfinite(a).all() and not torch.equal(a, b)
    last = agent.statistics.last
    for k in ('loss_total', 'loss_policy', 'loss_entropy', 'ratio', 'entropy', 'loss_value', 'loss_v', 'gradients_norm_policy',
              'gradients_norm_dynamics', 'gradients_norm_value', 'returns', 'advantages', 'episode_rewards'):
        assert k in last and np.isfinite(last[k]), k
    assert agent.statistics.stats['gradients_norm_dynamics']['step'] == 4 * 264      # 264 trainable tensors x 4 minibatches
    assert any(f.startswith('events') for _, _, fs in os.walk(tmp_path / 'logs') for f in fs)      # Summary -> TensorBoard
    # save() wrote the three models + config.json (rl/agents/agents.py:195-203); load() restores them
    for f in ('policy_net.npz', 'value_net.npz', 'dynamics_model.npz', 'config.json'):
        assert os.path.exists(tmp_path / 'w' / 't' / f), f
    before = net.dynamics.flat.clone()
    net.dynamics.flat.zero_()
    agent.load()
    assert torch.equal(net.dynamics.flat, before)


@pytest.mark.parametrize('dtype,tol', [('f32', 1e-4), ('bf16', 5e-2)])
def test_rollout_inference_matches_oracle_on_the_trained_agent(built_libs, tmp_path, dtype, tol):
    dyn, pol, val = _reference_style_checkpoint(str(tmp_path / 'w' / 'stage-s5'))
    agent = _agent(tmp_path, batch_size=8, dtype=dtype, name='stage-s5', load=True, load_full=True, log_mode=None)
    net = agent.network
    d64 = {k: v.double() for k, v in dyn.items()}; p64 = {k: v.double() for k, v in pol.items()}; v64 = {k: v.double() for k, v in val.items()}
    for B, seed in ((1, 5), (8, 6)):                          # rollout step (batch of one) and a vectorised batch
        obs = C.synthetic_obs(B, H, Wd, seed=seed)
        torch.manual_seed(123)
        action, mean, std, log_prob, value = net.predict({k: v.cuda() for k, v in obs.items()})
        x = model.dynamics_forward(d64, C.oracle_obs(obs), training=False)           # moving statistics (core/networks.py:206-208)
        out = model.policy_forward(p64, x, action.double().cpu(), training=False)
        vout = model.value_forward(v64, x, training=False)
        assert C.rel_max(mean, out['mean']) < tol and C.rel_max(std, out['std']) < tol
        assert C.rel_max(value, vout['value']) < tol
        lp = model.beta_log_prob(out['alpha'], out['beta'], action.double().cpu())
        assert (log_prob.double().cpu() - lp).abs().max().item() < max(tol, 1e-4) * max(1.0, lp.abs().max().item())
        assert ((action > 0) & (action < 1)).all()
    # predict_last_value: zeros for a terminal state (core/networks.py:171,215-216), the value head otherwise
    obs1 = {k: v.cuda() for k, v in C.synthetic_obs(1, H, Wd, seed=9).items()}
    assert torch.equal(net.predict_last_value(obs1, is_terminal=True), torch.zeros(1, 2))
    lv = net.predict_last_value(obs1, is_terminal=False)
    x = model.dynamics_forward(d64, C.oracle_obs({k: v.cpu() for k, v in obs1.items()}), training=False)
    assert C.rel_max(lv, model.value_forward(v64, x, training=False)['value']) < tol
    # inference must not touch the moving statistics
    got = net.engine.dyn_state.to_dict()
    assert all(torch.equal(got[k].cpu(), dyn[k]) for k in got)


def test_augmented_rollout_on_cuda(built_libs, tmp_path, monkeypatch):
    """`aug_intensity` > 0 (stages 4-5 of the curriculum, main.py:79,88): the preprocess closure augments on the device and
    the rollout / update run on the augmented frames."""
    monkeypatch.chdir(tmp_path)
    agent = _agent(tmp_path, batch_size=8, aug_intensity=1.0, name='aug')
    fn = agent.preprocess()
    states = [agent.env.reset() for _ in range(3)]
    batch = fn(states)                                       # list of observation dicts -> batched dict, image augmented
    img = batch['state_image']
    assert img.is_cuda and img.dtype == torch.float32 and tuple(img.shape) == (3, 4, H, Wd, 3)
    assert torch.isfinite(img).all() and img.min() >= 0.0 and img.max() <= 1.0 + 1e-6
    for b in range(3):                                       # per-sample min-max normalisation (cutout / dropout only add zeros)
        assert img[b].min() == 0.0 and img[b].max() > 0.99
    raw = torch.as_tensor(np.stack([s['image'] for s in states], 0)).float() / 255
    assert not torch.allclose(img.cpu(), raw, atol=1e-3)
    agent.learn(episodes=1, timesteps=17, save_every='end', close=False)
    assert agent.network.engine.adam_step == dict(dyn=4, pol=2, val=2)
    assert np.isfinite(agent.statistics.last['loss_total'])
