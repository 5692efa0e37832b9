"""Debug aid: tcgen05 weight-gradient kernel vs the mma.sync kernel on the same saved activations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-driving-rl-agent_b200')]
import torch
from tests import common as C
from cdra.engine import Engine
from cdra import _lib
lib = _lib.load()
B, H, W = int(os.environ.get('B', 8)), 90, 120
dyn, pol, val = C.trained_params(torch.float64)
eng = Engine(B, H, W, dtype='bf16', image_u8=True, device='cuda')
C.load_engine(eng, dyn, pol, val)
dev = lambda d: {k: v.cuda() for k, v in d.items()}
obs, bt = dev(C.synthetic_obs(B, H, W, seed=71)), dev(C.synthetic_batch(B, seed=72))
x = eng.dynamics_forward(obs)
eng.policy_head(x, bt['actions'], bt['logp_old'], bt['adv'], bt['true_speed'], bt['true_sim'], 0.2, 1.0)
grads = {}
for tc in (1, 0, 0):
    lib.cdra_debug_set(b'tc', tc)
    eng.dynamics_backward(obs, eng.d_x512)
    torch.cuda.synchronize()
    grads.setdefault(tc, []).append(eng.dyn.to_dict(eng.g_dyn.clone()))
for k, g in grads[0][0].items():
    if k.startswith('tower.') and ('.pw' in k or '.scpw' in k or 'head' in k) and g.abs().max().item() > 1e-12:
        print(f'{k:24s} tc-vs-mma {C.rel_l2(grads[1][0][k], g):.3e}   mma-vs-mma {C.rel_l2(grads[0][1][k], g):.3e}  max {g.abs().max().item():.2e}')
