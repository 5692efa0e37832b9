"""Pre-defined architectures that operate over 'time' — interface of the reference's core/architectures.py.

In the reference these functions instantiate Keras layers; here they validate the hyper-parameters and
return the sub-network *specification* that `core.networks.dynamics_layers` hands to the CUDA plan (the
kernels are specialised for the shipped agents' configuration, core/carla_agent.py:61-68)."""
from typing import Dict


def feature_net(inputs, time_horizon: int, units=32, num_layers=2, activation='relu', normalization=None) -> Dict:
    """core/architectures.py:9-27: per time slice Dense(units, activation) -> BatchNorm, `num_layers` times,
    weights shared across slices, BatchNorm statistics per slice."""
    act = getattr(activation, '__name__', activation)
    if units != 16 or num_layers != 2 or act not in ('relu6',) or normalization is not None:
        raise NotImplementedError(f'feature_net(units={units}, num_layers={num_layers}, activation={act}, '
                                  f'normalization={normalization}): libcdra implements the shipped configuration '
                                  '(units=16, num_layers=2, relu6, no input normalisation)')
    return dict(kind='feature_net', input=inputs, time_horizon=time_horizon, units=units, num_layers=num_layers)


def shufflenet_v2(inputs, time_horizon: int, g=1.0, leak=0.0, last_channels=1024) -> Dict:
    """core/architectures.py:30-173: time-shared ShuffleNet-v2 (stem 3x3/2 + maxpool, stages of 4/8/4 units,
    1x1 head conv, global average pool), applied to every one of `time_horizon` frames."""
    assert g in [0.5, 1.0, 1.5, 2.0]                                     # :31
    if g != 1.0 or leak != 0.0 or last_channels != 768:
        raise NotImplementedError(f'shufflenet_v2(g={g}, leak={leak}, last_channels={last_channels}): libcdra implements '
                                  'g=1.0 (116/232/464 channels), ReLU6, last_channels=768 (core/carla_agent.py:66)')
    return dict(kind='shufflenet_v2', input=inputs, time_horizon=time_horizon, g=g, leak=leak, last_channels=last_channels)
