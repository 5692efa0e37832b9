"""Networks of the CARLA agent — API of the reference's core/networks.py on top of libcdra.

`CARLANetwork(agent, control_policy, control_value, dynamics, update_dynamics)` keeps the reference's
constructor and attribute surface (`.dynamics`, `.policy`, `.old_policy`, `.value`, `exp_scale`, `last_value`,
`predict`, `predict_last_value`, `dynamics_predict(_train)`, `value_predict`, `update_old_policy`, `reset`,
`save_weights`, `load_weights(full)`), but the three "models" are views of flat fp32 arenas that the CUDA
kernels read and the fused clip+Adam kernel updates in place.
"""
import os
from typing import Dict, List

import numpy as np
import torch

from cdra import checkpoint
from cdra.engine import Engine
from cdra.init import init_arena
from core import architectures as nn
from rl import utils
from rl.networks import Network


# -------------------------------------------------------------------------------------------------
# -- SHARED NETWORK
# -------------------------------------------------------------------------------------------------
def linear_combination(inputs, units=32, normalization='batch', name=None):
    """core/networks.py:24-30: BatchNorm -> Dense(units, linear)."""
    if normalization != 'batch':
        raise NotImplementedError('linear_combination without batch normalisation is not built')
    return dict(kind='linear_combination', input=inputs, units=units, name=name)


def dynamics_layers(inputs: dict, time_horizon: int, **kwargs):
    """core/networks.py:37-56: the shared-network architecture; returns the specification of its last layer
    (consumed by the CUDA plan) after validating it against what libcdra implements."""
    image_out = nn.shufflenet_v2(inputs['state_image'], time_horizon, **kwargs.get('shufflenet', {}))
    road_out = nn.feature_net(inputs['state_road'], time_horizon, **kwargs.get('road', dict(normalization=None)))
    vehicle_out = nn.feature_net(inputs['state_vehicle'], time_horizon, **kwargs.get('vehicle', {}))
    navigation_out = nn.feature_net(inputs['state_navigation'], time_horizon, **kwargs.get('navigation', {}))
    args = kwargs.get('rnn')
    if dict(args) != dict(image=256, road=32, vehicle=32, navigation=32):
        raise NotImplementedError(f'rnn={args}: libcdra implements GRU sizes image=256, road=vehicle=navigation=32')
    dyn = kwargs.get('dynamics', {})
    if dyn.get('units', 32) != 512:
        raise NotImplementedError('dynamics units must be 512 (core/carla_agent.py:68)')
    dynamics_in = dict(kind='concatenate', inputs=[image_out, road_out, vehicle_out, navigation_out], rnn=dict(args))
    return linear_combination(dynamics_in, **dyn, name='dynamics-linear')


def control_branch(inputs: dict, units: int, num_layers: int, activation=None):
    """core/networks.py:59-66: num_layers x [BatchNorm -> Dense(units, swish6)]."""
    act = getattr(activation, '__name__', activation)
    if units != 320 or num_layers != 2 or act not in ('swish6', None):
        raise NotImplementedError(f'control_branch(units={units}, num_layers={num_layers}, activation={act}) is not built')
    return dict(kind='control_branch', input=inputs['dynamics'], units=units, num_layers=num_layers)


# -------------------------------------------------------------------------------------------------
class ArenaModel:
    """What the reference gets from a `tf.keras.Model`: variables, get/set_weights, save/load, summary —
    backed by a (trainable arena, state arena) pair of the engine."""

    def __init__(self, name, arena, state, order=None, kind='dynamics'):
        self.name, self.arena, self.state, self.kind = name, arena, state, kind
        self._order = order          # Keras `get_weights()` order = layers in creation order, each [trainable..., moving...]

    @property
    def flat(self):
        return self.arena.flat

    @property
    def trainable_variables(self) -> List[torch.Tensor]:
        return [self.arena.view(n) for n in self.arena.names]

    def variable_names(self):
        names = []
        seen = set()
        for n in self.arena.names + self.state.names:
            layer = n.rsplit('.', 1)[0]
            if layer not in seen:
                seen.add(layer)
                names += [m for m in self.arena.names if m.rsplit('.', 1)[0] == layer]
                names += [m for m in self.state.names if m.rsplit('.', 1)[0] == layer]
        return names

    def _view(self, n):
        return self.arena.view(n) if n in self.arena.index else self.state.view(n)

    def get_weights(self) -> List[np.ndarray]:
        return [self._view(n).detach().cpu().numpy().copy() for n in self.variable_names()]

    def set_weights(self, weights):
        names = self.variable_names()
        assert len(weights) == len(names)
        for n, w in zip(names, weights):
            self._view(n).copy_(torch.as_tensor(np.asarray(w), dtype=torch.float32))

    def count_params(self):
        return self.arena.size + self.state.size

    def _named_views(self):
        return [(n, self._view(n)) for n in self.variable_names()]

    def save_weights(self, filepath):
        checkpoint.save_model(filepath, self._named_views())

    def load_weights(self, filepath, by_name=False):
        """`<filepath>.npz` (written by save_weights) or the reference's TensorFlow checkpoint `<filepath>.index`
        (weights/stage-*/, core/networks.py:302-310)."""
        checkpoint.load_model(filepath, self.kind, self._named_views())

    def summary(self):
        print(f'Model: "{self.name}"')
        for n in self.variable_names():
            print(f'  {n:36s} {tuple(self._view(n).shape)}')
        print(f'Total params: {self.count_params():,}  (trainable {self.arena.size:,}, non-trainable {self.state.size:,})')


class OldPolicy(ArenaModel):
    """`old_policy`: a detached copy of the policy arena (core/networks.py:175-176); only read during rollouts."""

    def __init__(self, policy: ArenaModel):
        self.name, self.kind = 'PolicyNetwork-old', 'policy'
        self._flat = policy.arena.flat.clone()
        self._state_flat = policy.state.flat.clone()
        self.arena, self.state = policy.arena, policy.state

    @property
    def flat(self):
        return self._flat

    def _view(self, n):
        return self.arena.view(n, self._flat) if n in self.arena.index else self.state.view(n, self._state_flat)

    def copy_from(self, policy: ArenaModel):
        self._flat.copy_(policy.arena.flat)
        self._state_flat.copy_(policy.state.flat)


class CARLANetwork(Network):
    """The CARLAgent network (core/networks.py:147-310)."""
    ENGINE = Engine

    def __init__(self, agent, control_policy: dict, control_value: dict, dynamics: dict, update_dynamics=False,
                 device=None, dtype=None):
        super().__init__(agent)
        self.inputs = self._get_input_layers()
        self.inputs['action'] = (agent.num_actions,)
        self.time_horizon = agent.env.time_horizon
        if self.time_horizon != 4:
            raise NotImplementedError('libcdra is built for env.time_horizon = 4 (core/carla_env.py:26)')
        self.spec = dynamics_layers(self.inputs, time_horizon=self.time_horizon, **dynamics)      # validates `dynamics`
        self.intermediate_inputs = dict(dynamics=(512,), action=self.inputs['action'])
        control_branch(self.intermediate_inputs, **control_policy)
        cv = dict(control_value)
        self.exp_scale = cv.pop('exponent_scale', 6.0)                                             # :247-248
        cv.pop('components', 1)
        control_branch(self.intermediate_inputs, **cv)
        if self.exp_scale != 6.0 or agent.num_actions != 2:
            raise NotImplementedError('exponent_scale must be 6 and num_actions 2 (core/networks.py:169, core/carla_env.py:18)')
        for k, d in (('state_road', 9), ('state_vehicle', 4), ('state_navigation', 5)):
            if tuple(agent.state_spec[k]) != (d,):
                raise NotImplementedError(f'{k} must have {d} features (core/carla_env.py:20-27)')
        H, W, C = agent.state_spec['state_image']
        assert C == 3

        if device is None:
            device = f'cuda:{torch.cuda.current_device()}'
        self.device = torch.device(device)
        self.dtype = dtype or 'bf16'
        self.image_u8 = bool(getattr(agent.env, 'image_uint8', False))
        self.engine = self.ENGINE(agent.batch_size, H, W, dtype=self.dtype, image_u8=self.image_u8, device=self.device)
        self._siblings: Dict[int, Engine] = {agent.batch_size: self.engine}
        from cdra.parallel import GradSync
        self.sync = GradSync(self.engine)            # one rank per GPU under torchrun; a no-op on a single process
        self.grad_scale = self.sync.grad_scale
        seed = agent.seed if agent.seed is not None else 42
        init_arena(self.engine.dyn, self.engine.dyn_state, seed)
        init_arena(self.engine.pol, self.engine.pol_state, seed + 1)
        init_arena(self.engine.val, self.engine.val_state, seed + 2)
        self.sync.broadcast_parameters(0)             # replicas start from rank 0's parameters whatever their seeds

        self.dynamics = ArenaModel('Dynamics-Model', self.engine.dyn, self.engine.dyn_state)
        self.action_index = 0
        self.value = ArenaModel('Value-Network', self.engine.val, self.engine.val_state, kind='value')
        self.last_value = torch.zeros((1, 2), dtype=torch.float32)                                  # (base, exp), :171
        self.policy = ArenaModel('PolicyNetwork', self.engine.pol, self.engine.pol_state, kind='policy')
        self.old_policy = OldPolicy(self.policy)
        self.update_old_policy()

    # ------------------------------------------------------------------ batching helpers
    def engine_for(self, batch) -> Engine:
        if batch not in self._siblings:
            self._siblings[batch] = self.engine.sibling(batch)
        return self._siblings[batch]

    def gather_device(self, tensors: List[torch.Tensor], index: torch.Tensor) -> List[torch.Tensor]:
        """Minibatch gather (the tf.data slicing of rl/utils.py:365-393) with cdra_gather_rows: `tensors` are contiguous
        device tensors [N, ...], `index` an int64 device vector."""
        out = [torch.empty((index.numel(),) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device) for t in tensors]
        return self.engine.gather_rows_multi(tensors, index, out)            # one launch for the whole minibatch

    def gather(self, tensors: List[torch.Tensor], idx: np.ndarray) -> List[torch.Tensor]:
        index = torch.as_tensor(idx, dtype=torch.int64, device=self.device)
        return self.gather_device([(t if t.dim() > 1 else t.unsqueeze(-1)).to(self.device).contiguous() for t in tensors], index)

    def _obs(self, states: dict):
        img = states['state_image']
        if self.image_u8 and img.dtype != torch.uint8:
            # float frames (e.g. augmented ones, in [0, 1]) for a uint8 engine: quantise to the byte grid the stem reads
            img = (img.float().clamp(0.0, 1.0) * 255.0).round().to(torch.uint8) if img.is_floating_point() else img.to(torch.uint8)
        elif not self.image_u8 and img.dtype != torch.float32:
            img = img.float()
        obs = dict(state_image=img.contiguous().to(self.device))
        for k in ('state_road', 'state_vehicle', 'state_navigation'):
            obs[k] = states[k].float().contiguous().to(self.device)
        return obs

    # ------------------------------------------------------------------ reference API
    def dynamics_predict_train(self, inputs: dict):
        """core/networks.py:210-212 (training=True: batch statistics, moving averages updated)."""
        obs = self._obs(inputs)
        eng = self.engine_for(obs['state_image'].shape[0])
        return dict(dynamics=eng.dynamics_forward(obs, training=True), action=inputs.get('action'), _obs=obs, _engine=eng)

    def dynamics_predict(self, inputs: dict):
        """core/networks.py:206-208 (training=False: moving statistics)."""
        obs = self._obs(inputs)
        eng = self.engine_for(obs['state_image'].shape[0])
        return dict(dynamics=eng.dynamics_forward(obs, training=False), action=inputs.get('action'), _obs=obs, _engine=eng)

    def value_predict(self, inputs):
        eng = inputs['_engine']
        B = eng.B
        z = torch.zeros(B, 2, device=self.device)
        eng.value_head(inputs['dynamics'], z, z[:, :1].contiguous(), z[:, :1].contiguous(), training=False, backward=False)
        return eng.head_out.view(-1)[:B * 4].view(B, 4)[:, :2].clone()

    def predict(self, inputs):
        dynamics_inputs = self.data_for_dynamics(inputs)
        dynamics_output = self.dynamics_predict(dynamics_inputs)
        return self._predict(inputs=dynamics_output)

    def _predict(self, inputs):
        """core/networks.py:187-193: old policy (sampled action, mean, std, log_prob) + value."""
        eng = inputs['_engine']
        B = eng.B
        x = inputs['dynamics']
        z2, z1 = torch.full((B, 2), 0.5, device=self.device), torch.zeros(B, 1, device=self.device)
        cur = eng.pol.flat.clone()
        cur_state = eng.pol_state.flat.clone()
        eng.pol.flat.copy_(self.old_policy._flat); eng.pol_state.flat.copy_(self.old_policy._state_flat)
        try:
            eng.policy_head(x, z2, z2, z1.view(-1), z1, z1, training=False, backward=False)
        finally:
            eng.pol.flat.copy_(cur); eng.pol_state.flat.copy_(cur_state)
        ho = eng.head_out.view(B, 8)
        alpha, beta = ho[:, 0:2], ho[:, 2:4]
        dist = torch.distributions.Beta(alpha, beta)
        action = dist.sample().clamp(utils.EPSILON, 1.0 - utils.EPSILON)                          # _clip_actions :139-144
        log_prob = dist.log_prob(action)
        value = self.value_predict(inputs)
        self.action_index += 1
        return action, dist.mean, dist.stddev, log_prob, value

    def act(self, inputs):
        return self.predict(inputs)[0]

    def data_for_dynamics(self, inputs):
        """core/networks.py:195-204: append the last action (the dynamics model passes it through untouched)."""
        inputs = dict(inputs)
        memory = self.agent.memory
        n = len(memory) if memory is not None else 0
        inputs['action'] = torch.zeros((1, self.agent.num_actions)) if n == 0 else memory.last_action()
        return inputs

    def predict_last_value(self, state, is_terminal: bool, **kwargs):
        if is_terminal:
            return self.last_value
        dynamics_out = self.dynamics_predict(self.data_for_dynamics(state))
        return self.value_predict(dynamics_out)

    def reset(self):
        super().reset()
        self.action_index = 0

    def update_old_policy(self, weights=None):
        if weights is not None and not isinstance(weights, bool):
            if isinstance(weights, torch.Tensor):
                self.old_policy._flat.copy_(weights)
            else:
                self.old_policy.set_weights(weights)
        else:
            self.old_policy.copy_from(self.policy)

    def summary(self):
        print('==== Policy Network ====')
        self.policy.summary()
        print('\n==== Value Network ====')
        self.value.summary()
        print('\n==== Dynamics Model ====')
        self.dynamics.summary()

    def save_weights(self):
        self.policy.save_weights(filepath=self.agent.weights_path['policy'])
        self.value.save_weights(filepath=self.agent.weights_path['value'])
        self.dynamics.save_weights(filepath=self.agent.dynamics_path)

    def load_weights(self, full=True):
        if full:
            self.policy.load_weights(filepath=self.agent.weights_path['policy'], by_name=False)
            self.old_policy.copy_from(self.policy)
            self.value.load_weights(filepath=self.agent.weights_path['value'], by_name=False)
            self.dynamics.load_weights(filepath=self.agent.dynamics_path, by_name=False)
        else:
            self.dynamics.load_weights(filepath=self.agent.dynamics_path, by_name=False)
