"""What an agent expects of its network object (interface of the reference's rl/networks/networks.py:13-110).  The one
concrete network of this build is core.networks.CARLANetwork, whose numerics live in libcdra."""

_REQUIRED = ('predict', 'act', 'trainable_variables', 'set_weights', 'get_weights', 'load_weights', 'save_weights')


class Network:
    def __init__(self, agent):
        self.agent = agent

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        cls.missing_hooks = tuple(n for n in _REQUIRED if getattr(cls, n) is getattr(Network, n))

    def reset(self):
        """start of an episode (recurrent state, if any)"""

    def summary(self):
        """print the layer table (optional)"""

    def _get_input_layers(self, include_actions=False) -> dict:
        """name -> per-sample shape of every state (and optionally action) component; there are no Keras Input layers
        here (rl/networks/networks.py:47-66), the shapes parameterise the CUDA plan instead."""
        spec = dict(self.agent.state_spec)
        return {**spec, **self.agent.action_spec} if include_actions else spec


def _required(name):
    def hook(self, *args, **kwargs):
        raise NotImplementedError(f'{type(self).__name__} does not implement {name}()')
    hook.__name__ = name
    return hook


for _name in _REQUIRED:
    setattr(Network, _name, _required(_name))
