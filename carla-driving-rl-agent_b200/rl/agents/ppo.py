"""Proximal Policy Optimization agent: the algorithmic skeleton of the reference's rl/agents/ppo.py
(hyper-parameters, update loop, gradient application order, rollout loop, memory) with the numerics
delegated to libcdra through `agent.network` (core.networks.CARLANetwork)."""
import os
import random
import time
from typing import Union

import numpy as np
import torch

from rl import utils
from rl.agents.agents import Agent
from rl.parameters import DynamicParameter, LearningRateSchedule


_Schedule = Union[float, LearningRateSchedule, DynamicParameter]
# the schedules that live in config.json next to the weights (rl/agents/ppo.py:598-617), attribute name == config key
_SCHEDULED = ('policy_lr', 'value_lr', 'adv_scale', 'entropy_strength', 'clip_ratio')


def _every(spec, total, name):
    """`save_every` / `render_every` of learn(): False / None -> never, True -> every episode, 'end' -> once after the last."""
    if spec is False or spec is None:
        return total + 1
    if spec is True:
        return 1
    if spec == 'end':
        return total
    if name == 'save_every' and total % spec:
        raise AssertionError(f'episodes ({total}) must be a multiple of {name} ({spec})')
    return spec


class PPOAgent(Agent):
    def __init__(self, *args, policy_lr: _Schedule = 1e-3, gamma=0.99, lambda_=0.95, value_lr: _Schedule = 3e-4, load=False,
                 optimization_steps=(1, 1), name='ppo-agent', optimizer='adam', clip_norm=(1.0, 1.0), clip_ratio: _Schedule = 0.2,
                 seed_regularization=False, entropy_regularization: _Schedule = 0.0, network: Union[dict, object] = None,
                 update_frequency=1, polyak=1.0, repeat_action=1, advantage_scale: _Schedule = 2.0, **kwargs):
        # keyword names, defaults and the checks below: rl/agents/ppo.py:20-111
        if not 0.0 < polyak <= 1.0:
            raise AssertionError('polyak must lie in (0, 1]')
        if repeat_action < 1:
            raise AssertionError('repeat_action must be >= 1')
        if isinstance(clip_ratio, float) and clip_ratio < 0.0:
            raise AssertionError('clip_ratio must be >= 0')
        if not isinstance(network, dict) or 'network' not in network:
            raise ValueError('this build only implements the CARLANetwork family: pass network=dict(network=CARLANetwork, ...)')
        super().__init__(*args, name=name, **kwargs)

        self.memory: PPOMemory = None
        self.gamma, self.lambda_ = gamma, lambda_
        self.repeat_action, self.update_frequency = repeat_action, update_frequency
        self.polyak_coeff, self.should_polyak_average = polyak, polyak < 1.0
        self.optimization_steps = dict(zip(('policy', 'value'), optimization_steps))

        # every schedulable hyper-parameter is a DynamicParameter (constant, schedule or user-provided)
        for attr, value in (('policy_lr', policy_lr), ('value_lr', value_lr), ('adv_scale', advantage_scale),
                            ('entropy_strength', entropy_regularization), ('clip_ratio', clip_ratio)):
            setattr(self, attr, DynamicParameter.create(value=value))

        # re-seed with a fresh random seed before every minibatch when asked to (a regulariser of the reference)
        self.seed_regularization = self._reseed if seed_regularization else (lambda: None)
        self.seed_regularization()

        self._init_action_space()
        for label, what in (('state_spec', self.state_spec), ('action_shape', self.num_actions), ('distribution', self.distribution_type)):
            print(f'{label}:', what)
        self._init_gradient_clipping(clip_norm)

        spec = dict(network)
        self.network = spec.pop('network')(agent=self, **spec)
        self.policy_optimizer, self.value_optimizer = (utils.get_optimizer_by_name(optimizer, learning_rate=lr)
                                                       for lr in (self.policy_lr, self.value_lr))
        if load:
            self.load()

    def _reseed(self):
        self.set_random_seed(random.randint(a=0, b=2 ** 32 - 1))

    def _init_gradient_clipping(self, clip_norm):
        """rl/agents/ppo.py:114-146."""
        if clip_norm is None:
            self.should_clip_policy_grads = self.should_clip_value_grads = False
        elif isinstance(clip_norm, float):
            assert clip_norm > 0.0
            self.should_clip_policy_grads = self.should_clip_value_grads = True
            self.grad_norm_policy = self.grad_norm_value = clip_norm
        else:
            assert isinstance(clip_norm, tuple)
            for i, tag in ((0, 'policy'), (1, 'value')):
                if clip_norm[i] is None:
                    setattr(self, f'should_clip_{tag}_grads', False)
                else:
                    assert isinstance(clip_norm[i], float) and clip_norm[i] > 0.0
                    setattr(self, f'should_clip_{tag}_grads', True)
                    setattr(self, f'grad_norm_{tag}', clip_norm[i])

    def _init_action_space(self):
        """rl/agents/ppo.py:149-183 restricted to what the CUDA heads implement: bounded Box -> Beta."""
        space = self.env.action_space
        if utils.spaces.kind(space) != 'box' or not space.is_bounded():
            raise NotImplementedError('only bounded continuous (Beta) action spaces are built (CARLAEnv, core/carla_env.py:18)')
        self.num_actions = space.shape[0]
        self.distribution_type = 'beta'
        self.action_low = torch.as_tensor(space.low, dtype=torch.float32)
        self.action_high = torch.as_tensor(space.high, dtype=torch.float32)
        self.action_range = self.action_high - self.action_low
        self.convert_action = lambda a: (a.detach().cpu() * self.action_range + self.action_low)[0].numpy()

    def act(self, state, *args, **kwargs):
        return self.convert_action(self.network.act(inputs=state))

    def predict(self, state, *args, **kwargs):
        return self.network.predict(inputs=state)

    # ------------------------------------------------------------------ update (rl/agents/ppo.py:190-226)
    def update(self):
        """All policy minibatches, then all value minibatches, each `optimization_steps[...]` times; every minibatch is
        gradients -> (all-reduce) -> clip -> Adam on the device, nothing is read back here."""
        started = time.time()
        self.seed_regularization()
        # both iterators are built up front (the reference draws the value shuffle first)
        batches = dict(value=self.get_value_batches(), policy=self.get_policy_batches())
        phases = (('policy', self.get_policy_gradients, self.update_policy, self.policy_lr, 'loss_total'),
                  ('value', self.get_value_gradients, self.update_value, self.value_lr, 'loss_value'))
        for tag, gradients_of, apply, lr, loss_key in phases:
            for _ in range(self.optimization_steps[tag]):
                for minibatch in batches[tag]():
                    self.seed_regularization()
                    loss, grads = gradients_of(minibatch)
                    apply(grads)
                    self.log(**{loss_key: loss, f'lr_{tag}': lr.value, f'gradients_norm_{tag}': self._head_norms})

        enqueued = time.time()                           # host time to enqueue the update (the device runs behind it)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        print(f'Update took {round(time.time() - started, 3)}s (enqueued in {round(enqueued - started, 3)}s)')

    def get_policy_gradients(self, batch):
        raise NotImplementedError

    def get_value_gradients(self, batch):
        raise NotImplementedError

    def update_policy(self, gradients):
        return self.apply_policy_gradients(gradients), True

    def update_value(self, gradients):
        return self.apply_value_gradients(gradients), True

    def apply_policy_gradients(self, gradients, reduced=False):
        """rl/agents/ppo.py:238-252: per-tensor clip -> (old_policy <- policy) -> Adam [-> Polyak].  `reduced`: the caller
        already exchanged the gradients across the data-parallel ranks."""
        net = self.network
        clip = self.grad_norm_policy if self.should_clip_policy_grads else None
        if not reduced:
            net.sync.allreduce('pol')
        self._head_norms = net.engine.grad_norms('pol', net.grad_scale) if self.statistics.should_log else None
        if self.should_polyak_average:
            old = net.policy.flat.clone()
            net.update_old_policy(old)
            net.engine.clip_adam('pol', self.policy_lr(), clip, net.grad_scale)
            utils.polyak_averaging(net.policy.flat, old, alpha=self.polyak_coeff)
        else:
            net.update_old_policy()
            net.engine.clip_adam('pol', self.policy_lr(), clip, net.grad_scale)
        return gradients

    def apply_value_gradients(self, gradients, reduced=False):
        """rl/agents/ppo.py:264-275."""
        net = self.network
        clip = self.grad_norm_value if self.should_clip_value_grads else None
        if not reduced:
            net.sync.allreduce('val')
        self._head_norms = net.engine.grad_norms('val', net.grad_scale) if self.statistics.should_log else None
        if self.should_polyak_average:
            old = net.value.flat.clone()
            net.engine.clip_adam('val', self.value_lr(), clip, net.grad_scale)
            utils.polyak_averaging(net.value.flat, old, alpha=self.polyak_coeff)
        else:
            net.engine.clip_adam('val', self.value_lr(), clip, net.grad_scale)
        return gradients

    def value_batch_tensors(self):
        return self.memory.states, self.memory.returns

    def policy_batch_tensors(self):
        return self.memory.states, self.memory.advantages, self.memory.actions, self.memory.log_probabilities

    def _batches(self, tensors, **kw):
        """Callable that yields gathered minibatches (tuples shaped like `tensors`) -- the data_to_batches role
        (rl/utils.py:365-393).  The rollout tensors are made device-resident ONCE, the index lists travel in one copy,
        every minibatch is one `cdra_gather_rows` per tensor."""
        flat, tree = _flatten(tensors)
        n = flat[0].shape[0]
        index_lists = utils.index_batches(n, self.batch_size, drop_remainder=self.drop_batch_remainder, skip=self.skip_count,
                                          num_shards=self.obs_skipping, seed=self.seed, **kw)
        # data parallel: one gradient all-reduce per minibatch, so every rank runs the same number of them
        index_lists = index_lists[:self.network.sync.agree_min(len(index_lists))]
        dev = self.network.device
        flat = [(t if t.dim() > 1 else t.unsqueeze(-1)).to(dev).contiguous() for t in flat]
        sizes = [len(ix) for ix in index_lists]
        all_idx = torch.as_tensor(np.concatenate(index_lists) if index_lists else np.zeros(0, np.int64), dtype=torch.int64).to(dev)
        starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

        def gen():
            for j, m in enumerate(sizes):
                yield _unflatten(self.network.gather_device(flat, all_idx[starts[j]:starts[j] + m]), tree)
        return gen

    def get_value_batches(self):
        """rl/agents/ppo.py:285-289 (the value pass always shuffles)."""
        return self._batches(self.value_batch_tensors(), shuffle=True, shuffle_batches=False)

    def get_policy_batches(self):
        """rl/agents/ppo.py:291-296."""
        return self._batches(self.policy_batch_tensors(), shuffle=self.shuffle, shuffle_batches=self.shuffle_batches)

    # ------------------------------------------------------------------ rollout (rl/agents/ppo.py:464-568)
    def learn(self, episodes: int, timesteps: int, save_every: Union[bool, str, int] = False,
              render_every: Union[bool, str, int] = False, close=True):
        if episodes % self.update_frequency:
            raise AssertionError('episodes must be a multiple of update_frequency')
        save_every, render_every = _every(save_every, episodes, 'save_every'), _every(render_every, episodes, 'render_every')
        try:
            self.memory = self.get_memory()
            for episode in range(1, episodes + 1):
                reward = self._collect_episode(episode, timesteps, render=episode % render_every == 0)
                if episode % self.update_frequency == 0:
                    self.update()
                    self.memory.delete()
                    self.memory = self.get_memory()
                # (update_frequency > 1: the reference trims the bootstrap entries here, :551-553; update_index already did)
                self.log(episode_rewards=reward)
                self.write_summaries()
                if self.should_record:
                    self.record(episode)
                self.on_episode_end()
                if episode % save_every == 0:
                    self.save()
        finally:
            if close:
                print('closing...')
                self.env.close()

    def _observe(self, state, preprocess_fn):
        """environment observation -> the `state_*` tensor dict the network consumes"""
        if isinstance(state, dict):
            state = {f'state_{k}': v for k, v in state.items()}
        return utils.to_tensor(preprocess_fn(state))

    def _collect_episode(self, episode, timesteps, render=False):
        """One trajectory of at most `timesteps` decisions into `self.memory`, closed with the bootstrap value and its
        returns / advantages (the body of the reference's episode loop, :497-545).  Returns the episode reward."""
        self.seed_regularization()
        self.on_episode_start()
        preprocess_fn = self.preprocess()
        self.reset()
        state, total, started = self.env.reset(), 0.0, time.time()
        for t in range(1, timesteps + 1):
            if render:
                self.env.render()
            state = self._observe(state, preprocess_fn)
            action, mean, std, log_prob, value = self.predict(state)
            action_env = self.convert_action(action)
            for _ in range(self.repeat_action):              # the same action `repeat_action` times, rewards summed
                next_state, reward, done, _ = self.env.step(action_env)
                total += reward
                if done:
                    break
            self.log(actions=action, action_env=action_env, rewards=reward, distribution_mean=mean, distribution_std=std)
            self.memory.append(state, action, reward, value, log_prob)
            state = next_state
            if done or t == timesteps:
                print(f'Episode {episode} terminated after {t} timesteps in {round(time.time() - started, 3)}s '
                      f'with reward {round(total, 3)}.')
                self.log(timestep=t)
                last_value = self.network.predict_last_value(self._observe(state, preprocess_fn), timestep=(t + 1) / timesteps,
                                                             is_terminal=done)
                self.end_episode(last_value, append=self.update_frequency > 1)
                break
        return total

    def get_memory(self):
        return PPOMemory(state_spec=self.state_spec, num_actions=self.num_actions)

    def end_episode(self, last_value, append=False):
        """rl/agents/ppo.py:574-585: returns + GAE for the trajectory that just ended (cdra_gae)."""
        self.memory.end_trajectory(last_value)
        returns, values, advantages = self.memory.compute_returns_and_advantages(
            self.network.engine, self.gamma, self.lambda_, scale=self.adv_scale(), append=append)
        self.memory.update_index(append=append)
        self.log(returns=returns, advantages=advantages, values=values, advantage_scale=self.adv_scale.value,
                 returns_base=self.memory.returns[:, 0], returns_exp=self.memory.returns[:, 1],
                 values_base=self.memory.values[:, 0], values_exp=self.memory.values[:, 1],
                 advantages_normalized=self.memory.advantages)

    def summary(self):
        self.network.summary()

    def save_weights(self):
        print('saving weights...')
        self.network.save_weights()

    def load_weights(self):
        print('loading weights...')
        self.network.load_weights()

    def save_config(self):
        print('save config')
        self.update_config(**{key: getattr(self, key).serialize() for key in _SCHEDULED})
        super().save_config()

    def load_config(self):
        print('load config')
        super().load_config()
        for key in _SCHEDULED:
            getattr(self, key).load(config=self.config.get(key, {}))

    def reset(self):
        super().reset()
        self.network.reset()

    def on_episode_end(self):
        super().on_episode_end()
        for schedule in (self.policy_lr, self.value_lr, self.adv_scale):
            schedule.on_episode()


def _flatten(tree):
    flat, spec = [], []
    for item in tree:
        if isinstance(item, dict):
            keys = list(item.keys())
            spec.append(keys)
            flat.extend(item[k] for k in keys)
        else:
            spec.append(None)
            flat.append(item)
    return flat, spec


def _unflatten(flat, spec):
    out, i = [], 0
    for s in spec:
        if s is None:
            out.append(flat[i]); i += 1
        else:
            out.append({k: flat[i + j] for j, k in enumerate(s)}); i += len(s)
    return tuple(out)


class PPOMemory:
    """Recent memory used in PPOAgent (rl/agents/ppo.py:629-754) as PREALLOCATED DEVICE BUFFERS: every `append` is one
    row write (host -> device copy straight into the buffer, uint8 frames stay uint8), the update's minibatches are
    gathered from the buffers by `cdra_gather_rows`, returns / advantages are written by `cdra_gae`.  (The reference
    re-`tf.concat`s every tensor at every step -- O(N^2) copies -- and keeps fp32 frames on the host.)

    `num_envs` > 1 stores vectorised rollouts: one `append` carries the transition of every environment; rows are
    time-major (row = step * num_envs + env), trajectories are the columns."""

    def __init__(self, state_spec: dict, num_actions: int, device='cpu', capacity=256, num_envs=1, image_dtype=torch.float32):
        self.index = 0                              # first step of the trajectory being collected
        self.device = torch.device(device)
        self.simple_state = list(state_spec.keys()) == ['state']
        self.state_spec = state_spec
        self.num_actions = num_actions
        self.num_envs = int(num_envs)
        self.capacity = max(1, int(capacity))       # environment steps
        self.size = 0                               # environment steps stored
        self.image_dtype = image_dtype
        self._buf = None
        self._finished = False                      # end_trajectory has appended the bootstrap row
        self.returns = None
        self.advantages = None

    # ------------------------------------------------------------------ storage
    def _alloc(self, steps):
        E, dev = self.num_envs, self.device
        f = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)
        buf = dict(actions=f(steps * E, self.num_actions), log_probs=f(steps * E, self.num_actions),
                   values=f((steps + 1) * E, 2), rewards=f((steps + 1) * E), states={})
        for k, shape in self.state_spec.items():
            dtype = self.image_dtype if k.endswith('image') else torch.float32
            lead = () if self.simple_state else (getattr(self, 'time_horizon', None),)
            lead = tuple(x for x in lead if x)
            buf['states'][k] = torch.zeros((steps * E,) + lead + tuple(shape), dtype=dtype, device=dev)
        return buf

    def _ensure(self, steps):
        if self._buf is None:
            self._buf = self._alloc(self.capacity)
        if steps <= self.capacity:
            return
        cap = self.capacity
        while cap < steps:
            cap *= 2
        new, n, E = self._alloc(cap), self.size, self.num_envs
        for k in ('actions', 'log_probs'):
            new[k][:n * E].copy_(self._buf[k][:n * E])
        for k in ('values', 'rewards'):
            new[k][:(n + 1) * E].copy_(self._buf[k][:(n + 1) * E])
        for k, v in self._buf['states'].items():
            new['states'][k][:n * E].copy_(v[:n * E])
        self._buf, self.capacity = new, cap

    def __len__(self):
        return self.size * self.num_envs

    def delete(self):
        self._buf = None
        self.size = self.index = 0
        self.returns = self.advantages = None

    def _rows(self, x, width=None):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        return t.reshape(self.num_envs, -1) if width is None else t.reshape(self.num_envs, width)

    def append(self, state, action, reward, value, log_prob):
        """One transition per environment (rl/agents/ppo.py:678-690).  Tensors may live on the host (pinned memory makes
        the copies asynchronous) or on the device; state tensors carry a leading axis of `num_envs`."""
        assert not self._finished, 'append after end_trajectory: call update_index / delete first'
        self._ensure(self.size + 1)
        E, lo = self.num_envs, self.size * self.num_envs
        b = self._buf
        states = {'state': state} if self.simple_state else state
        for k, dst in b['states'].items():
            src = states[k] if isinstance(states[k], torch.Tensor) else torch.as_tensor(np.asarray(states[k]))
            if dst.dtype == torch.uint8 and src.is_floating_point():     # float frames in [0, 1] (augmented) -> the byte grid the stem reads
                src = (src.clamp(0.0, 1.0) * 255.0).round()
            dst[lo:lo + E].copy_(src.reshape(dst[lo:lo + E].shape), non_blocking=True)
        b['actions'][lo:lo + E].copy_(self._rows(action, self.num_actions), non_blocking=True)
        b['log_probs'][lo:lo + E].copy_(self._rows(log_prob, self.num_actions), non_blocking=True)
        b['values'][lo:lo + E].copy_(self._rows(value, 2), non_blocking=True)
        r = reward if isinstance(reward, torch.Tensor) else torch.as_tensor(np.asarray(reward, dtype=np.float32))
        b['rewards'][lo:lo + E].copy_(r.reshape(E), non_blocking=True)
        self.size += 1

    def last_action(self):
        E = self.num_envs
        return self._buf['actions'][(self.size - 1) * E: self.size * E]

    # ------------------------------------------------------------------ views the agent batches over
    @property
    def states(self):
        n = len(self)
        views = {k: v[:n] for k, v in self._buf['states'].items()}
        return views['state'] if self.simple_state else views

    @property
    def actions(self):
        return self._buf['actions'][:len(self)]

    @property
    def log_probabilities(self):
        return self._buf['log_probs'][:len(self)]

    @property
    def values(self):
        return self._buf['values'][:(self.size + (1 if self._finished else 0)) * self.num_envs]

    @property
    def rewards(self):
        return self._buf['rewards'][:(self.size + (1 if self._finished else 0)) * self.num_envs]

    # ------------------------------------------------------------------ returns / advantages
    def end_trajectory(self, last_value: torch.Tensor):
        """Adds the value of the terminal state (rl/agents/ppo.py:692-697): values gets (base, exp), rewards gets
        v_T = base * 10^exp as the bootstrap.  last_value: [num_envs, 2] (zeros for terminal states)."""
        self._ensure(self.size + 1)
        E, lo = self.num_envs, self.size * self.num_envs
        lv = torch.as_tensor(last_value, dtype=torch.float32).reshape(E, 2).to(self.device)
        self._buf['values'][lo:lo + E].copy_(lv)
        v_T = (lv[:, 0].double() * torch.pow(torch.tensor(10.0, dtype=torch.float64, device=self.device), lv[:, 1].double())).float()
        self._buf['rewards'][lo:lo + E].copy_(v_T)
        self._finished = True

    def compute_returns_and_advantages(self, engine, gamma: float, lambda_: float, scale=2.0, append=False):
        """compute_returns + compute_advantages (rl/agents/ppo.py:699-727) of the trajectory (per environment) that just
        ended, in one cdra_gae call; results are stored time-major like every other buffer."""
        assert self._finished
        E, n0, n1 = self.num_envs, self.index, self.size
        T = n1 - n0
        dev = engine.device
        r = self._buf['rewards'][n0 * E:n1 * E].view(T, E).t().contiguous().to(dev)
        vbe = self._buf['values'][n0 * E:n1 * E].view(T, E, 2).transpose(0, 1).contiguous().to(dev)
        last = self._buf['values'][n1 * E:(n1 + 1) * E].contiguous().to(dev)
        returns_be, adv = engine.gae(r, vbe, last, gamma, lambda_, scale)               # [E,T,2], [E,T]
        new_returns = returns_be.transpose(0, 1).reshape(T * E, 2).to(self.device)
        new_adv = adv.t().reshape(T * E).to(self.device)
        if (self.returns is None) or (not append):
            self.returns, self.advantages = new_returns, new_adv
        else:
            self.returns = torch.cat([self.returns, new_returns], 0)
            self.advantages = torch.cat([self.advantages, new_adv], 0)
        ten = torch.tensor(10.0, device=self.device)
        vals = self._buf['values'][n0 * E:(n1 + 1) * E]
        values = vals[:, 0] * torch.pow(ten, vals[:, 1])
        returns = new_returns[:, 0] * torch.pow(ten, new_returns[:, 1])
        return returns, values, new_adv

    def update_index(self, append=False):
        """rl/agents/ppo.py:729-733: the next trajectory starts after this one; with `append` (update_frequency > 1)
        the bootstrap row is dropped and collection continues."""
        self.index = self.size
        self._finished = False

    def serialize(self, episode: int, save_path: str):
        filename = f'trace-{episode}-{time.strftime("%Y%m%d-%H%M%S")}.npz'
        buffer = dict(reward=self.rewards.cpu().numpy(), action=self.actions.cpu().numpy(), value=self.values.cpu().numpy(),
                      log_prob=self.log_probabilities.cpu().numpy())
        if self.simple_state:
            buffer['state'] = self.states.cpu().numpy()
        else:
            for key, value in self.states.items():
                buffer[key] = value.cpu().numpy()
        np.savez_compressed(file=os.path.join(save_path, filename), **buffer)
        print(f'Traces "{filename}" saved.')
