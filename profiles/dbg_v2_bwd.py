"""Debug aid: v2 (bf16) policy pass gradients vs the fp64 oracle (trained weights): per-tensor rel L2 + cosine."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-driving-rl-agent_b200')]
import torch
from tests import common as C
from cdra.engine import Engine

B, H, W = 8, 90, 120
obs, bt = C.synthetic_obs(B, H, W, seed=41), C.synthetic_batch(B, seed=42)
dev = lambda d: {k: v.cuda() for k, v in d.items()}
dyn, pol, val = C.trained_params(torch.float64)
eng = Engine(B, H, W, dtype='bf16', image_u8=True, device='cuda')
C.load_engine(eng, dyn, pol, val)
sc = C.policy_step_engine(eng, dev(obs), dev(bt)).cpu()
torch.cuda.synchronize()
ref = C.policy_step_oracle(dyn, pol, obs, bt)
print('loss', sc[0].item(), ref['loss'].item())
mine = eng.dyn.to_dict(eng.g_dyn)
rows = []
for k, g in ref['g_dyn'].items():
    if g.abs().max().item() < 1e-9:
        continue
    m = mine[k].double().cpu().flatten(); g = g.flatten()
    cos = (m @ g / (m.norm() * g.norm() + 1e-30)).item()
    rows.append((k, C.rel_l2(mine[k], ref['g_dyn'][k]), cos, g.abs().max().item()))
for r in rows:
    if r[0].startswith('tower'):
        print(f'{r[0]:28s} rel_l2 {r[1]:.3f} cos {r[2]:.4f} max {r[3]:.2e}')
import statistics
tw = [r for r in rows if r[0].startswith('tower')]
print('median rel_l2', statistics.median(r[1] for r in tw), 'min cos', min(r[2] for r in tw), 'finite', torch.isfinite(eng.g_dyn).all().item())
