"""Debug aid: tensor-core stem (v2_stem.cuh) vs the legacy CUDA-core stem kernels on identical inputs (bf16 mode)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-driving-rl-agent_b200')]
import torch
from tests import common as C
from cdra.engine import Engine

B, H, W = int(os.environ.get('B', 8)), 90, 120
obs, bt = C.synthetic_obs(B, H, W, seed=41), C.synthetic_batch(B, seed=42)
dev = lambda d: {k: v.cuda() for k, v in d.items()}
dyn, pol, val = C.trained_params(torch.float64) if os.environ.get('TRAINED') else C.fresh_params(torch.float64)
eng = Engine(B, H, W, dtype='bf16', image_u8=True, device='cuda')
C.load_engine(eng, dyn, pol, val)
o = dev(obs)
sc = C.policy_step_engine(eng, o, dev(bt)).cpu()
torch.cuda.synchronize()
g_full = eng.dyn.to_dict(eng.g_dyn)
g_new = eng.dyn.to_dict(eng.debug_stem_backward(o, False))
g_old = eng.dyn.to_dict(eng.debug_stem_backward(o, True))
torch.cuda.synchronize()
for k in ('tower.stem.w', 'tower.stem.g', 'tower.stem.be', 'tower.stem.b'):
    print(f'{k:14s} new-vs-legacy rel_l2 {C.rel_l2(g_new[k], g_old[k]):.3e} rel_max {C.rel_max(g_new[k], g_old[k]):.3e}  '
          f'replay-vs-step {C.rel_l2(g_new[k], g_full[k]):.3e}  max|legacy| {g_old[k].abs().max().item():.3e}')
print('loss', sc[0].item())
