"""Pins the oracle against every artefact of the reference that constrains this path:
checkpoint structure of the 6 shipped agents, closed-form known answers, library-call identities."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ckpt, model, ppo, spec
from tests.common import GOLDEN, trained_params


def test_param_totals_match_shipped_checkpoints():
    idx = json.load(open(os.path.join(GOLDEN, 'ckpt_index.json')))
    known = json.load(open(os.path.join(GOLDEN, 'known_answers.json')))
    want = {}
    for name, pspec in (('dynamics_model', spec.dynamics_params()), ('policy_net', spec.head_params('policy')),
                        ('value_net', spec.head_params('value'))):
        want[name] = sum(spec.numel(s) for _, s, _ in pspec)
    assert want == known['totals']
    for stage, d in idx.items():
        assert d['totals'] == want, stage


def test_every_checkpoint_variable_has_an_oracle_counterpart():
    """shape multiset of the oracle inventory == shape multiset of the checkpoint index"""
    idx = json.load(open(os.path.join(GOLDEN, 'ckpt_index.json')))['stage-s5-curriculum']['variables']

    def norm(shape):
        s = [d for d in shape if d != 1] or [1]
        return tuple(s)

    for fname, pspec in (('dynamics_model', spec.dynamics_params()), ('policy_net', spec.head_params('policy')),
                         ('value_net', spec.head_params('value'))):
        a = sorted(norm(s) for _, s in idx[fname])
        b = sorted(norm(s) for _, s, _ in pspec)
        assert a == b, fname


def test_layer_order_mapping_is_total():
    assert len(ckpt.dynamics_layer_order()) == 130          # layer_with_weights-0..129
    assert sum(1 for _, k in ckpt.dynamics_layer_order() if k == 'bn') == 63
    tr, nt, st, ns = spec.split_layout(spec.dynamics_params())
    assert (len(tr), nt, ns) == (264, 2128450, 16564)
    for kind, n in (('policy', 270470), ('value', 269828)):
        tr, nt, st, ns = spec.split_layout(spec.head_params(kind))
        assert (len(tr), nt, ns) == (16, n, 1664)


def test_trained_fixture_loads_and_runs():
    dyn, pol, val = trained_params(torch.float32)
    assert set(dyn) == {n for n, _, _ in spec.dynamics_params()}
    B = 2
    obs = dict(state_image=torch.rand(B, 4, 90, 120, 3), state_road=torch.rand(B, 4, 9), state_vehicle=torch.rand(B, 4, 4),
               state_navigation=torch.rand(B, 4, 5))
    x = model.dynamics_forward(dyn, obs, training=False)
    out = model.policy_forward(pol, x, torch.rand(B, 2), training=False)
    v = model.value_forward(val, x, training=False)
    assert x.shape == (B, 512) and torch.isfinite(x).all()
    assert (out['alpha'] > 1.0).all() and (out['beta'] > 1.0).all()        # softplus + 1.01 (networks.py:133-134)
    assert (v['value'][:, 1] >= 0).all() and (v['value'][:, 1] <= 6).all()


def test_known_answers():
    known = json.load(open(os.path.join(GOLDEN, 'known_answers.json')))
    for x, base, e in known['decompose_number']:
        b, ee = ppo.decompose_number(x)
        assert ee == e and abs(float(b) - base) < 1e-6
    assert model.shuffle_perm(8) == known['shuffle_c8']
    x = torch.arange(8.0).view(1, 1, 1, 8)
    assert model.channel_shuffle(x).flatten().tolist() == [float(i) for i in known['shuffle_c8']]
    assert [list(s) for s in spec.spatial_sizes(90, 120)] == known['spatial_90x120']


def test_sp_norm_extremes():
    x = np.array([-3.0, -1.0, 0.0, 2.0, 5.0], np.float32)
    y = ppo.sp_norm(x)
    assert abs(y.max() - 5.0 / 5.001) < 1e-6 and abs(y.min() + 3.0 / 3.001) < 1e-6 and y[2] == 0.0   # rl/utils.py:344-349


def test_tf_same_padding_is_asymmetric():
    assert spec.same_pad(22, 3, 2) == (0, 1) and spec.same_pad(11, 3, 2) == (1, 1) and spec.same_pad(59, 3, 2) == (1, 1)
    x = torch.zeros(1, 4, 4, 1); x[0, 3, 3, 0] = 1.0
    w = torch.zeros(3, 3, 1); w[2, 2, 0] = 1.0          # bottom-right tap sees the padded cell after the input
    y = model.depthwise3x3(x, w, torch.zeros(1), 2)
    assert y.shape == (1, 2, 2, 1) and y.sum() == 0.0
    w = torch.zeros(3, 3, 1); w[1, 1, 0] = 1.0          # pad_before = 0: centre tap of out(1,1) reads in(3,3)
    assert model.depthwise3x3(x, w, torch.zeros(1), 2)[0, 1, 1, 0] == 1.0      # (symmetric padding would read in(2,2))


def test_gru_matches_torch_gru_with_reordered_gates():
    torch.manual_seed(0)
    B, D, U = 3, 5, 4
    k, r, b = torch.randn(D, 3 * U), torch.randn(U, 3 * U), torch.randn(2, 3 * U)
    xs = [torch.randn(B, D) for _ in range(4)]
    h = model.gru(xs, k, r, b)
    g = torch.nn.GRU(D, U, batch_first=True)
    perm = torch.cat([torch.arange(U, 2 * U), torch.arange(0, U), torch.arange(2 * U, 3 * U)])   # [z|r|h] -> [r|z|n]
    with torch.no_grad():
        g.weight_ih_l0.copy_(k.t()[perm]); g.weight_hh_l0.copy_(r.t()[perm])
        g.bias_ih_l0.copy_(b[0][perm]); g.bias_hh_l0.copy_(b[1][perm])
    out, hn = g(torch.stack(xs, 1))
    assert torch.allclose(h, hn[0], atol=1e-5)


def test_beta_math_against_torch_distributions():
    a, b = torch.tensor([1.5, 3.0, 20.0]), torch.tensor([2.5, 1.01, 7.0])
    x = torch.tensor([0.2, 0.9, 0.6])
    d = torch.distributions.Beta(a, b)
    assert torch.allclose(model.beta_log_prob(a, b, x), d.log_prob(x), atol=1e-5)
    assert torch.allclose(model.beta_entropy(a, b), d.entropy(), atol=1e-5)


def test_adam_is_keras_flavoured():
    p, g = torch.tensor([1.0]), torch.tensor([0.5])
    m, v = torch.zeros(1), torch.zeros(1)
    ppo.adam_step(p, g, m, v, 1, 0.1)
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(p.item() - (1.0 - lr_t * 0.05 / (np.sqrt(0.00025) + 1e-7))) < 1e-6


def test_clip_by_norm():
    g = torch.tensor([3.0, 4.0])
    assert torch.allclose(ppo.clip_by_norm(g, 1.0), g / 5.0)
    assert torch.allclose(ppo.clip_by_norm(g * 0.01, 1.0), g * 0.01)
