"""`core` package of the reference (core/__init__.py:2-4) for the PPO-update hot path.  `CARLAEnv` (the CARLA
simulator environment) is untouched reference code; it is re-exported only when the reference tree and its
simulator dependencies are importable."""
from core.carla_agent import CARLAgent, FakeCARLAEnvironment, SyntheticCARLAEnvironment, CARLAMemory
from core.networks import CARLANetwork
