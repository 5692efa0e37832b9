"""CARLAgent — the reference's core/carla_agent.py agent surface (constructor kwargs, update path, losses,
gradient application order, memory, fake environment) with the numerics executed by libcdra.

Not mirrored: `record` / `evaluate` (CARLA roll-outs) — they need the simulator.  `augment` / `preprocess` run the
reference's image augmentation on the device (`cdra_augment`).
"""
import os
from typing import Union

import numpy as np
import torch

from rl import utils, spaces
from rl.agents.ppo import PPOAgent, PPOMemory
from rl.parameters import DynamicParameter, LearningRateSchedule
from core.networks import CARLANetwork


def swish6(x):
    """rl/utils.py:420-421 (name is what the network spec validates against)."""
    return torch.minimum(x * torch.sigmoid(x), torch.full_like(x, 6.0))


def relu6(x):
    return torch.clamp(x, 0.0, 6.0)


def sample_beta_reparameterized(alpha, beta):
    """A Beta(alpha, beta) sample the way tfp.distributions.Beta draws it [lib] -- x = g1 / (g1 + g2), g1 ~ Gamma(alpha),
    g2 ~ Gamma(beta) -- with its pathwise derivatives (dx/dalpha, dx/dbeta) from the implicit-reparameterisation gradients
    of the gamma draws.  Returns (x [B,2] contiguous, jac [B,2,2] contiguous); x is NOT clipped (the kernel clips it like
    `_clip_actions`, core/networks.py:139-144, and stops the gradient where it did)."""
    g1, g2 = torch._standard_gamma(alpha), torch._standard_gamma(beta)
    s = g1 + g2
    x = g1 / s
    jac = torch.stack([g2 / (s * s) * torch._standard_gamma_grad(alpha, g1), -g1 / (s * s) * torch._standard_gamma_grad(beta, g2)], dim=-1)
    return x.contiguous(), jac.contiguous()


class FakeCARLAEnvironment:
    """A testing-only environment with the state- and action-space of a CARLA environment
    (core/carla_agent.py:26-52).  Like the reference's it only defines spaces; unlike the reference's it uses the
    real `CARLAEnv` shapes (time_horizon 4, 5 waypoints, 2 actions, core/carla_env.py:18-27) so that an agent built
    on it matches the shipped checkpoints, and it carries the `info_buffer` / `reset_info` that
    `CARLAgent.update` needs."""
    time_horizon = 4

    def __init__(self, image_shape=(90, 120, 3), image_uint8=False):
        self.image_uint8 = image_uint8
        self.action_space = spaces.Box(low=-1.0, high=1.0, shape=(2,))                       # CARLAEnv.ACTION, carla_env.py:18
        self.observation_space = spaces.Dict(
            road=spaces.Box(low=0.0, high=15.0, shape=(9,)), vehicle=spaces.Box(low=-1.0, high=1.0, shape=(4,)),
            navigation=spaces.Box(low=0.0, high=25.0, shape=(5,)), image=spaces.Box(low=0.0, high=1.0, shape=image_shape))
        self.info_buffer = dict(speed=[], similarity=[])

    def reset_info(self):
        for k in self.info_buffer:
            self.info_buffer[k] = []

    def seed(self, seed):
        pass

    def step(self, action):
        pass

    def reset(self):
        pass

    def render(self, mode='human'):
        pass

    def close(self):
        pass


class SyntheticCARLAEnvironment(FakeCARLAEnvironment):
    """FakeCARLAEnvironment that actually steps: observations / rewards / info drawn from the synthetic
    distributions of SURVEY §8(d).  Drives `CARLAgent.learn` end to end without a simulator."""

    def __init__(self, image_shape=(90, 120, 3), image_uint8=True, seed=0):
        super().__init__(image_shape, image_uint8)
        self.rng = np.random.RandomState(seed)
        self.image_shape = image_shape

    def _obs(self):
        r = self.rng
        img = r.randint(0, 256, size=(4,) + tuple(self.image_shape)).astype(np.uint8)
        if not self.image_uint8:
            img = img.astype(np.float32) / 255.0
        road = np.concatenate([(r.rand(4, 3) < 0.2).astype(np.float32), 0.3 + 0.6 * r.rand(4, 1), np.eye(5)[r.randint(0, 5, 4)]], 1)
        veh = np.concatenate([r.rand(4, 1) * 2 - 1, r.rand(4, 3)], 1)
        nav = np.sort(r.rand(4, 5) * 25, axis=1)
        return dict(image=img, road=road.astype(np.float32), vehicle=veh.astype(np.float32), navigation=nav.astype(np.float32))

    def reset(self):
        return self._obs()

    def step(self, action):
        self.info_buffer['speed'].append(float(self.rng.rand() * 30.0))
        self.info_buffer['similarity'].append(float(self.rng.rand() * 2 - 1))
        reward = float(np.clip(self.rng.randn() * 2 + 1, -10, 30))
        return self._obs(), reward, False, {}


# -------------------------------------------------------------------------------------------------
# -- Agent
# -------------------------------------------------------------------------------------------------
class CARLAgent(PPOAgent):
    DEFAULT_CONTROL = dict(units=320, num_layers=2, activation=swish6)                           # core/carla_agent.py:61-68
    DEFAULT_CONTROL_VALUE = dict(units=320, num_layers=2, activation=swish6)
    DEFAULT_DYNAMICS = dict(road=dict(units=16, num_layers=2, activation=relu6),
                            vehicle=dict(units=16, num_layers=2, activation=relu6),
                            navigation=dict(units=16, num_layers=2, activation=relu6),
                            shufflenet=dict(g=1.0, last_channels=768),
                            rnn=dict(image=256, road=32, vehicle=32, navigation=32),
                            dynamics=dict(units=512))

    def __init__(self, *args, aug_intensity=1.0, clip_norm=(1.0, 1.0, 1.0), name='carla', load_full=True, eta=0.0,
                 dynamics_lr: Union[float, LearningRateSchedule] = 1e-3, update_dynamics=True, delta=0.0, aux=1.0,
                 **kwargs):
        assert aug_intensity >= 0.0                                                              # :84
        network_spec = dict(kwargs.pop('network', {}))
        network_spec.setdefault('network', CARLANetwork)
        network_spec.setdefault('control_policy', self.DEFAULT_CONTROL)
        network_spec.setdefault('control_value', self.DEFAULT_CONTROL_VALUE)
        network_spec.setdefault('dynamics', self.DEFAULT_DYNAMICS)

        self.should_update_dynamics = update_dynamics
        self.reparameterized_actions = kwargs.pop('reparameterized_actions', True)     # False: stored-action style constant
        self.dynamics_path = os.path.join(kwargs.get('weights_dir', 'weights'), name, 'dynamics_model')
        self.load_full = load_full
        super().__init__(*args, name=name, network=network_spec, clip_norm=clip_norm, **kwargs)
        self.network: CARLANetwork = self.network
        self.aug_intensity = aug_intensity
        self.delta, self.eta, self.aux = delta, eta, aux                                         # stored, unused (:102-104)
        self.evaluation_path = utils.makedir(os.path.join(self.base_path, 'evaluation'))

        # the reference sets these flags but never clips the dynamics gradients (:109-117 vs :386-388)
        if isinstance(clip_norm, float):
            self.should_clip_dynamics_grads, self.grad_norm_dynamics = True, clip_norm
        elif isinstance(clip_norm[2], float):
            self.should_clip_dynamics_grads, self.grad_norm_dynamics = True, clip_norm[2]
        else:
            self.should_clip_dynamics_grads = False

        self.dynamics_lr = DynamicParameter.create(value=dynamics_lr)
        self.dynamics_lr.load(config=self.config.get('dynamics_lr', {}))
        self.dynamics_optimizer = utils.get_optimizer_by_name(name=kwargs.get('optimizer', 'adam'), learning_rate=dynamics_lr)

    # ------------------------------------------------------------------ update (core/carla_agent.py:129-145)
    def update(self):
        # under data parallelism the decision is collective (a rank that skipped would leave the others' all-reduce hanging)
        if self.network.sync.agree_min(len(self.memory)) < self.batch_size:
            print('[Not updated] memory too small!')
            self.env.reset_info()
            return
        super().update()
        try:
            actions = (self.memory.actions - 1.0) * 2.0 + 1.0
            self.log(action_throttle_or_brake=actions[:, 0], action_steer=actions[:, 1])
        except Exception:
            print('[update] unable to print actions')
        self.env.reset_info()

    def _aux_targets(self, n):
        """speed / 100 and similarity from `env.info_buffer` (core/carla_agent.py:328-329,338-347), truncated or zero-padded
        to the memory length.  The buffers may be python lists (the reference's CARLAEnv) or tensors (vectorised feeders)."""
        def column(x):
            t = x.detach().float() if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float32))
            return t.reshape(-1, 1)
        speed, similarity = column(self.env.info_buffer['speed']) / 100.0, column(self.env.info_buffer['similarity'])
        if speed.shape[0] >= n:                                                                  # :338-345
            speed, similarity = speed[:n], similarity[:n]
        else:
            pad = torch.zeros(n - speed.shape[0], 1, device=speed.device)
            speed, similarity = torch.cat([speed, pad], 0), torch.cat([similarity, pad], 0)
        return speed, similarity

    def policy_batch_tensors(self):
        """core/carla_agent.py:323-332."""
        states, advantages, actions, log_probabilities = super().policy_batch_tensors()
        states = dict(states)
        states['action'] = actions
        speed, similarity = self._aux_targets(actions.shape[0])
        return states, advantages, log_probabilities, speed, similarity

    def value_batch_tensors(self):
        """core/carla_agent.py:334-349."""
        states, returns = super().value_batch_tensors()
        states = dict(states)
        states['action'] = self.memory.actions
        speed, similarity = self._aux_targets(returns.shape[0])
        return states, returns, speed, similarity

    # ------------------------------------------------------------------ gradients (:351-388, :430-463)
    def _named_grads(self, arena, flat):
        # the per-tensor views of a gradient buffer never change: build them once per (arena, buffer), not per minibatch
        cache = self.__dict__.setdefault('_grad_views', {})
        key = (id(arena), flat.data_ptr())
        if key not in cache:
            cache[key] = [arena.view(n, flat) for n in arena.names]
        return cache[key]

    def get_policy_gradients(self, batch):
        states, advantages, log_probabilities, speed, similarity = batch
        dynamics_out = self.network.dynamics_predict_train(states)
        new_batch = (dynamics_out, advantages, log_probabilities, speed, similarity)
        loss = self.policy_objective(batch=new_batch)
        eng = dynamics_out['_engine']
        grads = dict(policy=self._named_grads(eng.pol, eng.g_pol))
        if self.should_update_dynamics:
            eng.dynamics_backward(dynamics_out['_obs'], eng.d_x512)
            grads['dynamics'] = self._named_grads(eng.dyn, eng.g_dyn)
            return loss, grads
        return loss, grads['policy']

    def apply_policy_gradients(self, gradients):
        if isinstance(gradients, dict):
            assert self.should_update_dynamics
            # data parallel: ONE collective for everything the pass produced (policy head + dynamics gradients are one
            # contiguous range of the engine's gradient buffer), then the three fused clip+Adam launches
            self.network.sync.allreduce_pass('policy')
            self.apply_dynamics_gradients(gradients=gradients['dynamics'], reduced=True)
            super().apply_policy_gradients(gradients=gradients['policy'], reduced=True)
            self.log(gradients_norm_dynamics=self._dyn_norms)
        else:
            super().apply_policy_gradients(gradients)

    def apply_dynamics_gradients(self, gradients, reduced=False):
        """Adam without clipping (core/carla_agent.py:386-388)."""
        net = self.network
        if not reduced:
            net.sync.allreduce('dyn')
        # what the reference logs as [tf.norm(g) for g in grads] (:382,461): one launch, stays on the device
        self._dyn_norms = net.engine.grad_norms('dyn', net.grad_scale) if self.statistics.should_log else None
        net.engine.clip_adam('dyn', self.dynamics_lr(), None, net.grad_scale)
        return gradients

    def policy_predict(self, inputs: dict) -> dict:
        raise NotImplementedError('the policy head is evaluated inside policy_objective (one fused kernel)')

    def policy_objective(self, batch):
        """core/carla_agent.py:394-428: clipped surrogate - entropy + auxiliary losses; forward + backward of the
        policy head in one fused call (gradients land in the engine's policy arena and d_x512)."""
        states, advantages, old_log_prob, true_speed, true_similarity = batch
        eng = states['_engine']
        x = states['dynamics']
        B = eng.B
        # PolicyNetwork.call evaluates log pi_new at a fresh, reparameterised sample of the NEW policy
        # (core/networks.py:97-100): draw it from the current parameters (one forward-only head call), together with its
        # pathwise derivatives, which the fused kernel folds into d loss / d (alpha, beta)
        with torch.no_grad():
            z2, z1 = torch.full((B, 2), 0.5, device=eng.device), torch.zeros(B, 1, device=eng.device)
            eng.policy_head(x, z2, z2, z1.view(-1), z1, z1, training=True, backward=False, update_moving=False)
            ho = eng.head_out.view(B, 8)
            actions_eval, actions_jac = sample_beta_reparameterized(ho[:, 0:2], ho[:, 2:4])
        sc = eng.policy_head(x, actions_eval, old_log_prob.contiguous(), advantages.reshape(-1).contiguous(),
                             true_speed.contiguous(), true_similarity.contiguous(), float(self.clip_ratio()),
                             float(self.entropy_strength()), training=True, grad_scale=1.0, backward=True,
                             actions_jac=actions_jac if self.reparameterized_actions else None)
        self.log(ratio=sc[5], log_prob=sc[6], entropy=sc[7], entropy_coeff=self.entropy_strength.value,
                 ratio_clip=self.clip_ratio.value, loss_speed_policy=sc[3], loss_policy=sc[1], loss_entropy=sc[2],
                 speed_pi=sc[8], loss_similarity_policy=sc[4], similarity_pi=sc[9])
        return sc[0].clone()

    def get_value_gradients(self, batch):
        states, returns, speed, similarity = batch
        dynamics_out = self.network.dynamics_predict_train(states)
        loss = self.value_objective(batch=(dynamics_out, returns, speed, similarity))
        eng = dynamics_out['_engine']
        grads = dict(value=self._named_grads(eng.val, eng.g_val))
        if self.should_update_dynamics:
            eng.dynamics_backward(dynamics_out['_obs'], eng.d_x512)
            grads['dynamics'] = self._named_grads(eng.dyn, eng.g_dyn)
            return loss, grads
        return loss, grads['value']

    def apply_value_gradients(self, gradients):
        if isinstance(gradients, dict):
            assert self.should_update_dynamics
            self.network.sync.allreduce_pass('value')
            self.apply_dynamics_gradients(gradients=gradients['dynamics'], reduced=True)
            super().apply_value_gradients(gradients=gradients['value'], reduced=True)
            self.log(gradients_norm_dynamics_v=self._dyn_norms)
        else:
            super().apply_value_gradients(gradients)

    def value_objective(self, batch):
        """core/carla_agent.py:469-486."""
        states, returns, true_speed, true_similarity = batch
        eng = states['_engine']
        sc = eng.value_head(states['dynamics'], returns.contiguous(), true_speed.contiguous(), true_similarity.contiguous(),
                            training=True, grad_scale=1.0, backward=True)
        self.log(speed_v=sc[4], similarity_v=sc[5], loss_v=sc[1], loss_speed_value=sc[2], loss_similarity_value=sc[3])
        return sc[0].clone()

    def get_memory(self, capacity=256, num_envs=1):
        return CARLAMemory(state_spec=self.state_spec, num_actions=self.num_actions, time_horizon=self.env.time_horizon,
                           device=self.network.device, capacity=capacity, num_envs=num_envs,
                           image_dtype=torch.uint8 if self.network.image_u8 else torch.float32)

    def preprocess(self):
        """Augmentation function used during the reinforcement learning phase (core/carla_agent.py:523-525)."""
        return self.augment()

    def augment(self):
        """Augmentation closure of the reference (core/carla_agent.py:527-579) on the DEVICE: `prepare` batches the list of
        observation dicts, then -- when `aug_intensity` > 0 -- one `cdra_augment` call applies colour jitter, blur,
        salt & pepper, gaussian noise, per-sample min-max normalisation, cutout and coarse dropout with the reference's
        chance gates (cdra/augment.py draws the per-call scalars; the image stays on the GPU as float32 in [0, 1])."""
        alpha = float(self.aug_intensity)
        rng = np.random.default_rng(self.seed)
        device = self.network.device

        def prepare(state):
            if isinstance(state, list):
                state = {k: np.stack([np.asarray(s[k]) for s in state], 0) for k in state[0]}
                state = {f'state_{k}': v for k, v in state.items()}
            return state

        def augment_fn(states):
            state = prepare(states)
            if alpha <= 0.0:
                return state
            from cdra import augment as A
            image = state['state_image']
            image = image if isinstance(image, torch.Tensor) else torch.as_tensor(np.asarray(image))
            if image.dtype != torch.uint8:
                image = image.float()
            image = image.to(device).contiguous()
            group = int(image.shape[1]) if image.dim() == 5 else int(image.shape[0]) if image.dim() == 4 else 1     # one sample = its time_horizon frames
            params, mask = A.draw_params(rng, alpha, group=group)
            state = dict(state)
            state['state_image'] = A.augment(image, params, mask)
            return state

        return augment_fn

    def load_weights(self):
        print('loading weights...')
        self.network.load_weights(full=self.load_full)


class CARLAMemory(PPOMemory):
    """core/carla_agent.py:586-596: states carry a `time_horizon` axis ([N, 4, H, W, 3] frames, uint8 on the device when the
    environment delivers uint8)."""

    def __init__(self, state_spec: dict, num_actions: int, time_horizon: int, device='cpu', capacity=256, num_envs=1,
                 image_dtype=torch.float32):
        self.time_horizon = time_horizon
        super().__init__(state_spec, num_actions, device=device, capacity=capacity, num_envs=num_envs, image_dtype=image_dtype)
