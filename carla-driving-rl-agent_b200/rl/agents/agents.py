"""Host-side base of every agent: what `core/learning.py` and `CARLAgent` expect from `rl.agents.Agent`
(reference interface: rl/agents/agents.py:14-203 -- constructor keywords, the attribute names below, the
`weights/<name>/{policy_net,value_net,config.json}` layout and the hook methods).  Only the interface is shared with the
reference; the PPO update behind it runs in libcdra."""
import json
import os
import random

import numpy as np
import torch

from rl import utils

# hooks a concrete agent has to provide / may override
_ABSTRACT = ('act', 'predict', 'update', 'learn', 'get_memory', 'summary', 'load_weights', 'save_weights')
_OPTIONAL = ('record', 'reset', 'on_episode_start', 'on_episode_end')


def _resolve_env(env):
    """an environment instance, or a gym id to instantiate"""
    if not isinstance(env, str):
        return env
    import gym
    return gym.make(env)


class Agent:
    def __init__(self, env, batch_size: int, seed=None, weights_dir='weights', name='agent', log_mode='summary',
                 drop_batch_remainder=False, skip_data=0, consider_obs_every=1, evaluation_dir='evaluation',
                 shuffle_batches=False, shuffle=True, traces_dir=None, summary_keys=None):
        self.env = _resolve_env(env)
        self.batch_size = batch_size
        self.seed = None
        self.set_random_seed(seed)

        # flat {name: (shape, dtype, bounds)} views of the gym spaces
        self.state_spec, self.action_spec = (utils.space_to_flat_spec(space=space, name=label) for space, label in
                                             ((self.env.observation_space, 'state'), (self.env.action_space, 'action')))
        # minibatch iteration options
        self.drop_batch_remainder, self.shuffle_batches, self.shuffle = drop_batch_remainder, shuffle_batches, shuffle
        self.skip_count, self.obs_skipping = skip_data, consider_obs_every

        # on-disk layout: weights/<name>/..., evaluation/<name>/, optional traces/<name>/
        root = os.path.join(weights_dir, name)
        self.base_path = root
        self.weights_path = {net: os.path.join(root, f'{net}_net') for net in ('policy', 'value')}
        self.config_path = os.path.join(root, 'config.json')
        self.evaluation_path = utils.makedir(os.path.join(evaluation_dir, name))
        self.should_record = isinstance(traces_dir, str)
        if self.should_record:
            self.traces_dir = utils.makedir(traces_dir, name)

        self.config = {}
        self.statistics = utils.Summary(mode=log_mode, name=name, keys=summary_keys)

    # ------------------------------------------------------------------ seeding (agents.py:61-72, torch in place of tensorflow)
    def set_random_seed(self, seed):
        if seed is None:
            return
        if not 0 <= seed < 2 ** 32:
            raise AssertionError(f'seed {seed} outside [0, 2^32)')
        for seeder in (torch.manual_seed, np.random.seed, random.seed, getattr(self.env, 'seed', None)):
            if seeder is not None:
                seeder(seed)
        self.seed = seed
        print(f'Random seed {seed} set.')

    # ------------------------------------------------------------------ logging
    def preprocess(self):
        return lambda x: x

    def log(self, **kwargs):
        self.statistics.log(**kwargs)

    def write_summaries(self):
        try:
            self.statistics.write_summaries()
        except Exception:                      # a failed TensorBoard write must not end a training run
            print('[write_summaries] error.')

    # ------------------------------------------------------------------ config.json next to the weights
    def update_config(self, **kwargs):
        self.config.update(kwargs)

    def load_config(self):
        with open(self.config_path) as f:
            self.config = json.load(f)
        print('config loaded.')
        print(self.config)

    def save_config(self):
        utils.makedir(self.base_path)
        with open(self.config_path, 'w') as f:
            json.dump(self.config, f)
        print('config saved.')

    def load(self):
        self.load_weights()
        self.load_config()

    def save(self):
        self.save_weights()
        self.save_config()


def _abstract(name):
    def method(self, *args, **kwargs):
        raise NotImplementedError(f'{type(self).__name__}.{name}')
    method.__name__ = name
    return method


def _noop(name):
    def method(self, *args, **kwargs):
        return None
    method.__name__ = name
    return method


for _n in _ABSTRACT:
    setattr(Agent, _n, _abstract(_n))
for _n in _OPTIONAL:
    setattr(Agent, _n, _noop(_n))
