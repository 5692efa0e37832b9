"""cProfile of CARLAgent.update() on a synthetic rollout (host-side cost per SGD step of the reference-facing API)."""
import cProfile, io, os, pstats, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'carla-driving-rl-agent_b200'))
import torch
from core import CARLAgent, SyntheticCARLAEnvironment
bs, H, W, steps = 512, 90, 120, 6
env = SyntheticCARLAEnvironment(image_shape=(H, W, 3), image_uint8=True, seed=0)
tmp = tempfile.mkdtemp()
agent = CARLAgent(env, batch_size=bs, name='prof', weights_dir=tmp + '/w', evaluation_dir=tmp + '/e', seed=42, skip_data=0,
                  drop_batch_remainder=True, shuffle=True, shuffle_batches=False, log_mode='log', aug_intensity=0.0,
                  network=dict(dtype='bf16', device='cuda'))
g = torch.Generator().manual_seed(1)
mem = agent.get_memory(capacity=steps, num_envs=bs)
def fill():
    mem.delete(); agent.memory = mem
    for _ in range(steps):
        st = dict(state_image=torch.randint(0, 256, (bs, 4, H, W, 3), dtype=torch.uint8, generator=g), state_road=torch.rand(bs, 4, 9, generator=g),
                  state_vehicle=torch.rand(bs, 4, 4, generator=g), state_navigation=torch.rand(bs, 4, 5, generator=g))
        mem.append(st, torch.rand(bs, 2, generator=g).clamp(1e-4, 1 - 1e-4), torch.randn(bs, generator=g), torch.rand(bs, 2, generator=g), torch.randn(bs, 2, generator=g))
    env.info_buffer = dict(speed=torch.rand(steps * bs).cuda() * 30, similarity=torch.rand(steps * bs).cuda())
    agent.end_episode(torch.rand(bs, 2))
fill(); agent.update()
fill()
pr = cProfile.Profile(); pr.enable(); agent.update(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(40); print(s.getvalue()[:6000])
