"""`rl` package of the reference (rl/__init__.py), restricted to the PPO-update hot path.

The simulator side (`rl.environments.carla`, augmentations) is untouched reference code and is imported
lazily only when present: `import rl` must work on a box without carla / pygame / gym."""
from rl import utils
from rl import spaces
from rl.parameters import DynamicParameter
from rl.agents import Agent, PPOAgent
from rl.agents.ppo import PPOMemory
