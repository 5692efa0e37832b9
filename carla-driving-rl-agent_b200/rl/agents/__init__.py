from rl.agents.agents import Agent
from rl.agents.ppo import PPOAgent
