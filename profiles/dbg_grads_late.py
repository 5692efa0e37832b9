"""bf16 perf mode vs fp32 parity mode (engine vs engine, trained weights): gradients of the layers nearest the loss, where
the bf16 rounding noise is still small, so a wrong stencil tap / tile index shows up as an O(1) error."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'carla-driving-rl-agent_b200')
from cdra.engine import Engine
from tests import common as C
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 90, 120
e32 = Engine(B, H, W, dtype='f32', image_u8=True, device='cuda')
e16 = Engine(B, H, W, dtype='bf16', image_u8=True, device='cuda')
dyn, pol, val = C.trained_params(torch.float32)
C.load_engine(e32, dyn, pol, val); C.load_engine(e16, dyn, pol, val)
dev = lambda d: {k: v.cuda() for k, v in d.items()}
obs, bt = dev(C.synthetic_obs(B, H, W, seed=41)), dev(C.synthetic_batch(B, seed=42))
C.policy_step_engine(e32, obs, bt); C.policy_step_engine(e16, obs, bt)
g16, g32 = e16.dyn.to_dict(e16.g_dyn), e32.dyn.to_dict(e32.g_dyn)
worst = 0.0
for k in g32:
    if k.startswith(('tower.s3', 'tower.head', 'tower.s2.u7', 'tower.s2.u6')) and k.endswith('.w') and g32[k].abs().max() > 1e-12:
        e = C.rel_l2(g16[k], g32[k]); worst = max(worst, e)
        print(f'{k:24s} rel_l2 {e:.3e}  cos {torch.nn.functional.cosine_similarity(g16[k].flatten(), g32[k].flatten(), dim=0).item():.5f}')
print('worst', worst)
