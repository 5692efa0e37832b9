"""`Network` interface of the reference (rl/networks/networks.py:13-110): what an agent expects of its network."""


class Network:
    def __init__(self, agent):
        self.agent = agent

    def predict(self, *args, **kwargs):
        raise NotImplementedError

    def act(self, *args, **kwargs):
        raise NotImplementedError

    def reset(self):
        pass

    def trainable_variables(self):
        raise NotImplementedError

    def set_weights(self, weights):
        raise NotImplementedError

    def get_weights(self):
        raise NotImplementedError

    def load_weights(self):
        raise NotImplementedError

    def save_weights(self):
        raise NotImplementedError

    def summary(self):
        pass

    def _get_input_layers(self, include_actions=False) -> dict:
        """name -> per-sample shape of every state component (rl/networks/networks.py:47-66); there are no Keras
        Input layers here, the shapes parameterise the CUDA plan instead."""
        layers = dict(self.agent.state_spec)
        if include_actions:
            layers.update(self.agent.action_spec)
        return layers
