"""Checkpoint import (reference: core/networks.py:297-310, rl/agents/agents.py:195-203): the product's
TensorFlow-bundle reader against (a) the reference's shipped checkpoints where they are available, cross-checked with the
oracle's independent reader, (b) bundles re-emitted from the golden weights by the test-side writer; and the agent-level
`CARLAgent(load=True, load_full=...)` path on them."""
import os

import numpy as np
import pytest
import torch

from oracle import ckpt, spec
from tests import common as C
from tests.golden import tf_bundle_writer as W

REF = '/root/reference/weights'
H, Wd = 42, 58


@pytest.fixture(autouse=True)
def _cpu_logic_build(monkeypatch, built_libs):
    from core.networks import CARLANetwork
    from tests.emu.engine import EmuEngine
    monkeypatch.setattr(CARLANetwork, 'ENGINE', EmuEngine)


def _write_reference_style_checkpoint(folder, dyn, pol, val):
    os.makedirs(folder, exist_ok=True)
    for fname, params, order in (('dynamics_model', dyn, ckpt.dynamics_layer_order()), ('policy_net', pol, ckpt.head_layer_order('policy')),
                                 ('value_net', val, ckpt.head_layer_order('value'))):
        W.write_bundle(os.path.join(folder, fname), W.keras_checkpoint_tensors({k: v.numpy() for k, v in params.items()}, order))


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkpoints not present on this machine')
@pytest.mark.parametrize('stage', ['stage-s5-curriculum', 'stage-s2'])
def test_product_reader_on_the_shipped_checkpoints(stage):
    from cdra import checkpoint
    want = ckpt.load_reference_checkpoint(os.path.join(REF, stage))
    for (fname, model, pspec), ref in zip((('dynamics_model', 'dynamics', spec.dynamics_params()), ('policy_net', 'policy', spec.head_params('policy')),
                                           ('value_net', 'value', spec.head_params('value'))), want):
        names = [n for n, _, _ in pspec]
        got = checkpoint.bundle_to_arena_dict(os.path.join(REF, stage, fname), model, names)
        assert set(got) == set(names)
        for n, shape, _ in pspec:
            assert tuple(got[n].shape) == tuple(shape), n
            assert np.array_equal(got[n], ref[n]), n
    total = sum(v.size for v in checkpoint.bundle_to_arena_dict(os.path.join(REF, stage, 'dynamics_model'), 'dynamics',
                                                                [n for n, _, _ in spec.dynamics_params()]).values())
    assert total == 2_145_014


def test_bundle_writer_reader_round_trip(tmp_path):
    from cdra import checkpoint
    dyn, pol, val = C.trained_params(torch.float32)
    _write_reference_style_checkpoint(str(tmp_path), dyn, pol, val)
    b = checkpoint.TensorBundle(str(tmp_path / 'dynamics_model'))
    assert len(b.keras_layers()) == 130                                      # 130 weighted layers
    got = checkpoint.bundle_to_arena_dict(str(tmp_path / 'dynamics_model'), 'dynamics', list(dyn))
    for k, v in dyn.items():
        assert np.array_equal(got[k], v.numpy()), k
    with pytest.raises(ValueError):                                          # a policy file is not a dynamics model
        checkpoint.bundle_to_arena_dict(str(tmp_path / 'policy_net'), 'dynamics', list(dyn))
    with pytest.raises(FileNotFoundError):
        checkpoint.load_model(str(tmp_path / 'nothing'), 'value', [])


@pytest.mark.parametrize('load_full', [True, False])
def test_agent_loads_reference_style_checkpoint(tmp_path, load_full):
    """CARLAgent(load=True): weights/<name>/{policy_net,value_net,dynamics_model} in the reference's format + config.json
    (rl/agents/ppo.py:601-616); load_full=False restores only the dynamics model (core/networks.py:302-310)."""
    import json
    from core import CARLAgent, FakeCARLAEnvironment
    dyn, pol, val = C.trained_params(torch.float32)
    folder = tmp_path / 'w' / 'stage'
    _write_reference_style_checkpoint(str(folder), dyn, pol, val)
    with open(folder / 'config.json', 'w') as f:
        json.dump(dict(policy_lr=dict(step=7)), f)
    env = FakeCARLAEnvironment(image_shape=(H, Wd, 3), image_uint8=True)
    agent = CARLAgent(env, batch_size=2, name='stage', weights_dir=str(tmp_path / 'w'), evaluation_dir=str(tmp_path / 'e'), seed=3,
                      load=True, load_full=load_full, log_mode=None, network=dict(device='cpu', dtype='f32'))
    net = agent.network
    got = net.engine.dyn.to_dict(); got.update(net.engine.dyn_state.to_dict())
    for k, v in dyn.items():
        assert torch.equal(got[k], v), k
    gp = net.engine.pol.to_dict(); gp.update(net.engine.pol_state.to_dict())
    same = all(torch.equal(gp[k], v) for k, v in pol.items())
    assert same == load_full
    if load_full:
        assert torch.equal(net.old_policy.flat, net.policy.flat)              # old_policy <- policy after loading (:305)
        gv = net.engine.val.to_dict()
        assert all(torch.equal(gv[k], v) for k, v in val.items() if k in gv)
    # save -> load round trip through the library's own format
    agent.save()
    before = net.engine.dyn.flat.clone()
    net.engine.dyn.flat.zero_()
    agent.load()
    assert torch.equal(net.engine.dyn.flat, before)
