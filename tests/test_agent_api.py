"""The reference-facing Python surface (`core.CARLAgent`, `core.CARLANetwork`, `rl.*`) drives the CUDA path
with the reference's call sequence (rl/agents/ppo.py:190-226).  CPU run: the logic-check build stands in for
the GPU library, at a tiny image size."""
import numpy as np
import pytest
import torch

H, W = 42, 58


@pytest.fixture(autouse=True)
def _cpu_logic_build(monkeypatch, built_libs):
    """bind the networks to the CPU logic-check build of the library (tests/emu); the product has no such switch"""
    from core.networks import CARLANetwork
    from tests.emu.engine import EmuEngine
    monkeypatch.setattr(CARLANetwork, 'ENGINE', EmuEngine)


def _agent(tmp_path, batch_size=2, **kw):
    from core import CARLAgent, SyntheticCARLAEnvironment
    env = SyntheticCARLAEnvironment(image_shape=(H, W, 3), image_uint8=True, seed=1)
    base = dict(seed=7, skip_data=1, drop_batch_remainder=True, log_mode='summary', policy_lr=3e-4, value_lr=3e-4,
                dynamics_lr=3e-4, entropy_regularization=1.0, clip_ratio=0.2, gamma=0.9999, lambda_=0.999)
    base.update(kw)
    return CARLAgent(env, batch_size=batch_size, name='t', weights_dir=str(tmp_path / 'w'), evaluation_dir=str(tmp_path / 'e'),
                     network=dict(device='cpu', dtype='f32'), **base)


def test_api_surface_and_defaults(built_libs, tmp_path):
    from core import CARLAgent, CARLANetwork, FakeCARLAEnvironment
    from rl import PPOAgent, PPOMemory, DynamicParameter, utils
    agent = _agent(tmp_path)
    assert isinstance(agent, PPOAgent) and isinstance(agent.network, CARLANetwork)
    for attr in ('update', 'learn', 'predict', 'save', 'load', 'summary', 'get_memory', 'preprocess', 'policy_objective',
                 'value_objective', 'get_policy_gradients', 'get_value_gradients', 'apply_policy_gradients',
                 'apply_value_gradients', 'apply_dynamics_gradients', 'policy_batch_tensors', 'value_batch_tensors'):
        assert callable(getattr(agent, attr)), attr
    assert agent.state_spec == dict(state_image=(H, W, 3), state_navigation=(5,), state_road=(9,), state_vehicle=(4,))
    assert agent.num_actions == 2 and agent.distribution_type == 'beta'
    net = agent.network
    assert net.exp_scale == 6.0 and tuple(net.last_value.shape) == (1, 2)
    assert net.dynamics.count_params() == 2145014 and net.policy.count_params() == 272134 and net.value.count_params() == 271492
    assert len(net.dynamics.trainable_variables) == 264 and len(net.policy.trainable_variables) == 16
    w = net.policy.get_weights()
    net.policy.set_weights([x * 0 + 1 for x in w])
    assert float(net.policy.flat.min()) == 1.0
    net.policy.set_weights(w)
    # constructor validation mirrors the reference's asserts (rl/agents/ppo.py:33-34, core/carla_agent.py:84)
    with pytest.raises(AssertionError):
        _agent(tmp_path, polyak=0.0)
    with pytest.raises(AssertionError):
        _agent(tmp_path, aug_intensity=-1.0)
    with pytest.raises(ValueError):
        utils.get_optimizer_by_name('lion')
    assert utils.decompose_number(2.34)[1] == 1.0
    # save / load round trip keeps the 3-file + config.json layout (core/networks.py:297-310, agents.py:195-203)
    agent.save()
    before = net.dynamics.flat.clone()
    net.dynamics.flat.zero_()
    agent.load()
    assert torch.equal(net.dynamics.flat, before)


def test_update_follows_reference_sequence(built_libs, tmp_path):
    agent = _agent(tmp_path)
    env, net = agent.env, agent.network
    agent.memory = agent.get_memory()
    rng = np.random.RandomState(0)
    obs = env.reset()
    for t in range(3):                                   # skip_data=1 drops the first transition -> one minibatch of 2
        state = {f'state_{k}': torch.as_tensor(v).unsqueeze(0) for k, v in obs.items()}
        action = torch.rand(1, 2)
        obs, reward, done, _ = env.step(action.numpy()[0])
        agent.memory.append(state, action, reward, torch.tensor([[rng.rand() * 2 - 1, rng.rand() * 3]]), torch.randn(1, 2) * 0.3)
    agent.end_episode(torch.tensor([[0.3, 1.0]]))
    assert tuple(agent.memory.returns.shape) == (3, 2) and tuple(agent.memory.advantages.shape) == (3,)
    assert float(agent.memory.advantages.abs().max()) <= 2.0 + 1e-5          # sp-norm * scale 2 (rl/utils.py:344-349)
    p0, v0, d0 = net.policy.flat.clone(), net.value.flat.clone(), net.dynamics.flat.clone()
    old0 = net.old_policy.flat.clone()
    agent.update()
    assert not torch.equal(net.policy.flat, p0) and not torch.equal(net.value.flat, v0) and not torch.equal(net.dynamics.flat, d0)
    assert torch.equal(net.old_policy.flat, p0) and torch.equal(old0, p0)   # old <- policy before the Adam step (ppo.py:249)
    assert net.engine.adam_step == dict(dyn=2, pol=1, val=1)                # the trunk steps once per pass
    # first Adam step: every coordinate moves by ~lr (Keras epsilon 1e-7), no matter the gradient scale
    assert abs(float((net.value.flat - v0).abs().max()) - 3e-4) < 2e-5
    assert env.info_buffer == dict(speed=[], similarity=[])                 # reset_info (core/carla_agent.py:145)
    keys = agent.statistics.stats
    for k in ('loss_total', 'loss_policy', 'loss_entropy', 'ratio', 'loss_value', 'loss_v', 'gradients_norm_policy',
              'gradients_norm_dynamics', 'gradients_norm_value'):
        assert k in keys, k
    # too-small memory: soft skip (core/carla_agent.py:130-133)
    agent.memory = agent.get_memory()
    agent.update()


def test_index_batches_semantics():
    from rl import utils
    b = utils.index_batches(10, 4, skip=1, drop_remainder=True, shuffle=False)
    assert [x.tolist() for x in b] == [[1, 2, 3, 4], [5, 6, 7, 8]]
    b = utils.index_batches(10, 4, skip=1, drop_remainder=False, shuffle=False)
    assert b[-1].tolist() == [9]
    b = utils.index_batches(9, 4, num_shards=2, shuffle=False)
    assert np.concatenate(b).tolist() == [0, 2, 4, 6, 8, 1, 3, 5, 7]       # shard(2,0) ++ shard(2,1) (rl/utils.py:374-382)
    b = utils.index_batches(64, 8, shuffle=True, seed=3, drop_remainder=True)
    flat = np.concatenate(b)
    assert sorted(flat.tolist()) == list(range(64))
    # tf.data shuffle(buffer_size=8) is local: item v enters the buffer at step v and the first pop happens at step 8,
    # so v can never be emitted before output position v - 8
    assert all(pos >= int(v) - 8 for pos, v in enumerate(flat)) and flat.tolist() != list(range(64))


def test_dynamic_parameters():
    from rl.parameters import DynamicParameter, ConstantParameter, StepDecay, ExponentialDecay, PolynomialDecay
    p = DynamicParameter.create(0.5)
    assert isinstance(p, ConstantParameter) and p() == 0.5 and p.serialize() == {}
    s = StepDecay(1.0, decay_steps=2, decay_rate=0.5, min_value=0.2)
    vals = []
    for _ in range(6):
        vals.append(s()); s.on_episode()
    assert vals == [1.0, 1.0, 0.5, 0.5, 0.25, 0.25]
    s.load(dict(step=10)); assert s() == 0.2
    e = ExponentialDecay(1.0, decay_steps=1, decay_rate=0.9)
    e.step = 2
    assert abs(e() - 0.81) < 1e-12
    q = PolynomialDecay(1.0, 0.0, decay_steps=4)
    q.step = 2
    assert abs(q() - 0.5) < 1e-12


def test_vectorised_device_memory(built_libs, tmp_path):
    """the rollout memory as preallocated buffers: vectorised appends (one transition per environment),
    growth, time-major rows, per-environment returns / advantages equal to the single-trajectory oracle
    (rl/agents/ppo.py:692-727), and an update() over it"""
    from oracle import ppo
    E, T = 2, 3
    agent = _agent(tmp_path, batch_size=E * T, skip_data=0)
    mem = agent.get_memory(capacity=2, num_envs=E)                      # forces a growth step
    rng = np.random.RandomState(4)
    rew = (rng.randn(T, E) * 2 + 1).astype(np.float32)
    val = np.stack([rng.rand(T, E) * 2 - 1, np.floor(rng.rand(T, E) * 4)], -1).astype(np.float32)
    img = rng.randint(0, 256, size=(T, E, 4, H, W, 3)).astype(np.uint8)
    for t in range(T):
        state = dict(state_image=torch.from_numpy(img[t]), state_road=torch.rand(E, 4, 9), state_vehicle=torch.rand(E, 4, 4),
                     state_navigation=torch.rand(E, 4, 5))
        mem.append(state, torch.rand(E, 2), torch.from_numpy(rew[t]), torch.from_numpy(val[t]), torch.randn(E, 2) * 0.3)
    assert len(mem) == T * E and mem.capacity >= T and mem.states['state_image'].dtype == torch.uint8
    assert torch.equal(mem.states['state_image'].view(T, E, 4, H, W, 3)[2, 1], torch.from_numpy(img[2, 1]))     # time-major rows
    last = np.array([[0.5, 1.0], [0.0, 0.0]], np.float32)
    agent.memory = mem
    agent.end_episode(torch.from_numpy(last))
    ret, adv = mem.returns.view(T, E, 2).numpy(), mem.advantages.view(T, E).numpy()
    for e in range(E):
        r_ref, a_ref, _ = ppo.end_trajectory(rew[:, e], val[:, e], last[e], agent.gamma, agent.lambda_, 2.0)
        assert np.array_equal(ret[:, e, 1], r_ref[:, 1]) and np.abs(ret[:, e, 0] - r_ref[:, 0]).max() < 2e-7
        assert np.abs(adv[:, e] - a_ref).max() < 1e-6
    agent.env.info_buffer = dict(speed=torch.rand(T * E) * 30, similarity=torch.rand(T * E) * 2 - 1)         # tensors are accepted too
    p0 = agent.network.policy.flat.clone()
    agent.update()
    assert agent.network.engine.adam_step == dict(dyn=2, pol=1, val=1) and not torch.equal(agent.network.policy.flat, p0)
