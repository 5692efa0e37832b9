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


class PPOAgent(Agent):
    def __init__(self, *args, policy_lr: Union[float, LearningRateSchedule, DynamicParameter] = 1e-3, gamma=0.99,
                 lambda_=0.95, value_lr: Union[float, LearningRateSchedule, DynamicParameter] = 3e-4, load=False,
                 optimization_steps=(1, 1), name='ppo-agent', optimizer='adam', clip_norm=(1.0, 1.0),
                 clip_ratio: Union[float, LearningRateSchedule, DynamicParameter] = 0.2, seed_regularization=False,
                 entropy_regularization: Union[float, LearningRateSchedule, DynamicParameter] = 0.0,
                 network: Union[dict, object] = None, update_frequency=1, polyak=1.0, repeat_action=1,
                 advantage_scale: Union[float, LearningRateSchedule, DynamicParameter] = 2.0, **kwargs):
        assert 0.0 < polyak <= 1.0                              # rl/agents/ppo.py:33-34
        assert repeat_action >= 1
        super().__init__(*args, name=name, **kwargs)

        self.memory: PPOMemory = None
        self.gamma = gamma
        self.lambda_ = lambda_
        self.repeat_action = repeat_action
        self.adv_scale = DynamicParameter.create(value=advantage_scale)

        if seed_regularization:                                  # :44-52
            def _seed_regularization():
                self.set_random_seed(random.randint(a=0, b=2 ** 32 - 1))
            self.seed_regularization = _seed_regularization
            self.seed_regularization()
        else:
            self.seed_regularization = lambda: None

        self.entropy_strength = DynamicParameter.create(value=entropy_regularization)
        if isinstance(clip_ratio, float):
            assert clip_ratio >= 0.0
        self.clip_ratio = DynamicParameter.create(value=clip_ratio)

        self._init_action_space()
        print('state_spec:', self.state_spec)
        print('action_shape:', self.num_actions)
        print('distribution:', self.distribution_type)
        self._init_gradient_clipping(clip_norm)

        self.weights_path = dict(policy=os.path.join(self.base_path, 'policy_net'),
                                 value=os.path.join(self.base_path, 'value_net'))
        if not isinstance(network, dict) or 'network' not in network:
            raise ValueError('this build only implements the CARLANetwork family: pass network=dict(network=CARLANetwork, ...)')
        network = dict(network)
        network_class = network.pop('network')
        self.network = network_class(agent=self, **network)

        self.update_frequency = update_frequency
        self.policy_lr = DynamicParameter.create(value=policy_lr)
        self.value_lr = DynamicParameter.create(value=value_lr)
        self.optimization_steps = dict(policy=optimization_steps[0], value=optimization_steps[1])
        self.policy_optimizer = utils.get_optimizer_by_name(optimizer, learning_rate=self.policy_lr)
        self.value_optimizer = utils.get_optimizer_by_name(optimizer, learning_rate=self.value_lr)
        self.should_polyak_average = polyak < 1.0
        self.polyak_coeff = polyak
        if load:
            self.load()

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
        t0 = time.time()
        self.seed_regularization()
        value_batches = self.get_value_batches()
        policy_batches = self.get_policy_batches()

        for _ in range(self.optimization_steps['policy']):
            for data_batch in policy_batches():
                self.seed_regularization()
                total_loss, policy_grads = self.get_policy_gradients(data_batch)
                self.update_policy(policy_grads)
                self.log(loss_total=total_loss, lr_policy=self.policy_lr.value, gradients_norm_policy=self._head_norms)

        for _ in range(self.optimization_steps['value']):
            for data_batch in value_batches():
                self.seed_regularization()
                value_loss, value_grads = self.get_value_gradients(data_batch)
                self.update_value(value_grads)
                self.log(loss_value=value_loss, lr_value=self.value_lr.value, gradients_norm_value=self._head_norms)

        if torch.cuda.is_available():
            torch.cuda.synchronize()
        print(f'Update took {round(time.time() - t0, 3)}s')

    def get_policy_gradients(self, batch):
        raise NotImplementedError

    def get_value_gradients(self, batch):
        raise NotImplementedError

    def update_policy(self, gradients):
        return self.apply_policy_gradients(gradients), True

    def update_value(self, gradients):
        return self.apply_value_gradients(gradients), True

    def apply_policy_gradients(self, gradients):
        """rl/agents/ppo.py:238-252: per-tensor clip -> (old_policy <- policy) -> Adam [-> Polyak]."""
        net = self.network
        clip = self.grad_norm_policy if self.should_clip_policy_grads else None
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

    def apply_value_gradients(self, gradients):
        """rl/agents/ppo.py:264-275."""
        net = self.network
        clip = self.grad_norm_value if self.should_clip_value_grads else None
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
        """Callable that yields gathered minibatches (tuples shaped like `tensors`) — the data_to_batches role."""
        flat, tree = _flatten(tensors)
        n = flat[0].shape[0]
        index_lists = utils.index_batches(n, self.batch_size, drop_remainder=self.drop_batch_remainder, skip=self.skip_count,
                                          num_shards=self.obs_skipping, seed=self.seed, **kw)
        # data parallel: one gradient all-reduce per minibatch, so every rank runs the same number of them
        index_lists = index_lists[:self.network.sync.agree_min(len(index_lists))]

        def gen():
            for idx in index_lists:
                yield _unflatten(self.network.gather(flat, idx), tree)
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
        assert episodes % self.update_frequency == 0
        if (save_every is False) or (save_every is None):
            save_every = episodes + 1
        elif save_every is True:
            save_every = 1
        elif save_every == 'end':
            save_every = episodes
        else:
            assert episodes % save_every == 0
        if render_every is False:
            render_every = episodes + 1
        elif render_every is True:
            render_every = 1
        try:
            self.memory = self.get_memory()
            for episode in range(1, episodes + 1):
                self.seed_regularization()
                self.on_episode_start()
                preprocess_fn = self.preprocess()
                self.reset()
                state = self.env.reset()
                episode_reward = 0.0
                t0 = time.time()
                render = episode % render_every == 0
                for t in range(1, timesteps + 1):
                    if render:
                        self.env.render()
                    if isinstance(state, dict):
                        state = {f'state_{k}': v for k, v in state.items()}
                    state = utils.to_tensor(preprocess_fn(state))
                    action, mean, std, log_prob, value = self.predict(state)
                    action_env = self.convert_action(action)
                    for _ in range(self.repeat_action):
                        next_state, reward, done, _ = self.env.step(action_env)
                        episode_reward += reward
                        if done:
                            break
                    self.log(actions=action, action_env=action_env, rewards=reward, distribution_mean=mean, distribution_std=std)
                    self.memory.append(state, action, reward, value, log_prob)
                    state = next_state
                    if done or (t == timesteps):
                        print(f'Episode {episode} terminated after {t} timesteps in {round((time.time() - t0), 3)}s ' +
                              f'with reward {round(episode_reward, 3)}.')
                        self.log(timestep=t)
                        if isinstance(state, dict):
                            state = {f'state_{k}': v for k, v in state.items()}
                        state = utils.to_tensor(preprocess_fn(state))
                        last_value = self.network.predict_last_value(state, timestep=(t + 1) / timesteps, is_terminal=done)
                        self.end_episode(last_value, append=self.update_frequency > 1)
                        break
                if episode % self.update_frequency == 0:
                    self.update()
                    self.memory.delete()
                    self.memory = self.get_memory()
                elif self.update_frequency > 1:
                    self.memory.rewards = self.memory.rewards[:-1]
                    self.memory.values = self.memory.values[:-1]
                self.log(episode_rewards=episode_reward)
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
        self.update_config(policy_lr=self.policy_lr.serialize(), value_lr=self.value_lr.serialize(),
                           adv_scale=self.adv_scale.serialize(), entropy_strength=self.entropy_strength.serialize(),
                           clip_ratio=self.clip_ratio.serialize())
        super().save_config()

    def load_config(self):
        print('load config')
        super().load_config()
        self.policy_lr.load(config=self.config.get('policy_lr', {}))
        self.value_lr.load(config=self.config.get('value_lr', {}))
        self.adv_scale.load(config=self.config.get('adv_scale', {}))
        self.entropy_strength.load(config=self.config.get('entropy_strength', {}))
        self.clip_ratio.load(config=self.config.get('clip_ratio', {}))

    def reset(self):
        super().reset()
        self.network.reset()

    def on_episode_end(self):
        super().on_episode_end()
        self.policy_lr.on_episode()
        self.value_lr.on_episode()
        self.adv_scale.on_episode()


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
    """Recent memory used in PPOAgent (rl/agents/ppo.py:629-754).  Transitions are appended to python lists and
    stacked once per trajectory (the reference re-`tf.concat`s every tensor at every step)."""

    def __init__(self, state_spec: dict, num_actions: int, device='cpu'):
        self.index = 0
        self.device = torch.device(device)
        self.simple_state = list(state_spec.keys()) == ['state']
        self.state_spec = state_spec
        self._states = [] if self.simple_state else {k: [] for k in state_spec}
        self._actions, self._log_probs, self._values, self._rewards = [], [], [], []
        self.num_actions = num_actions
        self.states = None if self.simple_state else {}
        self.rewards = torch.zeros(0)
        self.values = torch.zeros(0, 2)
        self.actions = torch.zeros(0, num_actions)
        self.log_probabilities = torch.zeros(0, num_actions)
        self.returns = None
        self.advantages = None

    def __len__(self):
        return len(self._actions) if self._actions else self.actions.shape[0]

    def delete(self):
        self._states = self._actions = self._log_probs = self._values = self._rewards = None
        self.states = self.rewards = self.values = self.actions = self.log_probabilities = self.returns = self.advantages = None

    def append(self, state, action, reward, value, log_prob):
        if self.simple_state:
            self._states.append(state)
        else:
            assert isinstance(state, dict)
            for k, v in state.items():
                self._states[k].append(v)
        self._actions.append(torch.as_tensor(action, dtype=torch.float32).reshape(1, -1))
        self._rewards.append(float(reward))
        self._values.append(torch.as_tensor(value, dtype=torch.float32).reshape(1, 2))
        self._log_probs.append(torch.as_tensor(log_prob, dtype=torch.float32).reshape(1, -1))

    def _materialise(self):
        cat = lambda xs: torch.cat([torch.as_tensor(x) for x in xs], dim=0).to(self.device)
        if self.simple_state:
            self.states = cat(self._states)
        else:
            self.states = {k: cat(v).contiguous() for k, v in self._states.items()}
        self.actions = cat(self._actions)
        self.log_probabilities = cat(self._log_probs)

    def end_trajectory(self, last_value: torch.Tensor):
        """Adds the value of the terminal state (rl/agents/ppo.py:692-697)."""
        self._materialise()
        last_value = torch.as_tensor(last_value, dtype=torch.float32).reshape(1, 2).cpu()
        self._last_value = last_value
        self.values = torch.cat([torch.cat(self._values, 0).cpu(), last_value], 0)
        v_T = (last_value[:, 0].double() * torch.pow(torch.tensor(10.0, dtype=torch.float64), last_value[:, 1].double())).float()
        self.rewards = torch.cat([torch.tensor(self._rewards, dtype=torch.float32), v_T], 0)

    def compute_returns_and_advantages(self, engine, gamma: float, lambda_: float, scale=2.0, append=False):
        """compute_returns + compute_advantages (rl/agents/ppo.py:699-727) in one cdra_gae call."""
        dev = engine.device
        r = self.rewards[self.index:-1].reshape(1, -1).contiguous().to(dev)
        vbe = self.values[self.index:-1].reshape(1, -1, 2).contiguous().to(dev)
        last = self.values[-1:].reshape(1, 2).contiguous().to(dev)
        returns_be, adv = engine.gae(r, vbe, last, gamma, lambda_, scale)
        new_returns, new_adv = returns_be[0].to(self.device), adv[0].to(self.device)
        if (self.returns is None) or (not append):
            self.returns, self.advantages = new_returns, new_adv
        else:
            self.returns = torch.cat([self.returns, new_returns], 0)
            self.advantages = torch.cat([self.advantages, new_adv], 0)
        values = self.values[self.index:, 0] * torch.pow(torch.tensor(10.0), self.values[self.index:, 1])
        returns = new_returns[:, 0] * torch.pow(torch.tensor(10.0, device=new_returns.device), new_returns[:, 1])
        return returns, values, new_adv

    def update_index(self, append=False):
        self.index = self.rewards.shape[0] - 1 if append else self.rewards.shape[0]

    def serialize(self, episode: int, save_path: str):
        filename = f'trace-{episode}-{time.strftime("%Y%m%d-%H%M%S")}.npz'
        buffer = dict(reward=self.rewards.numpy(), action=self.actions.cpu().numpy(), value=self.values.numpy(),
                      log_prob=self.log_probabilities.cpu().numpy())
        if self.simple_state:
            buffer['state'] = self.states.cpu().numpy()
        else:
            for key, value in self.states.items():
                buffer[key] = value.cpu().numpy()
        np.savez_compressed(file=os.path.join(save_path, filename), **buffer)
        print(f'Traces "{filename}" saved.')
