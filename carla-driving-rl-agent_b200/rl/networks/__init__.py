from rl.networks.networks import Network
