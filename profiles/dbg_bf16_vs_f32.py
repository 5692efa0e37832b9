"""bf16 perf mode vs fp32 parity mode of the same kernels at a realistic minibatch (engine vs engine, on the GPU)."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'carla-driving-rl-agent_b200')
from cdra.engine import Engine
from cdra.init import init_engine
from tests import common as C
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 90, 120
e32 = Engine(B, H, W, dtype='f32', image_u8=True, device='cuda'); init_engine(e32, 42)
e16 = Engine(B, H, W, dtype='bf16', image_u8=True, device='cuda'); init_engine(e16, 42)
if len(sys.argv) > 2 and sys.argv[2] == 'trained':      # the reference's shipped stage-s5-curriculum agent
    dyn, pol, val = C.trained_params(torch.float32)
    C.load_engine(e32, dyn, pol, val); C.load_engine(e16, dyn, pol, val)
dev = lambda d: {k: v.cuda() for k, v in d.items()}
obs, bt = dev(C.synthetic_obs(B, H, W, seed=41)), dev(C.synthetic_batch(B, seed=42))
s32 = C.policy_step_engine(e32, obs, bt).cpu(); s16 = C.policy_step_engine(e16, obs, bt).cpu()
for k in ['tower.stem','tower.pool','tower.s1.u0.pw1','tower.s1.u3.dw','tower.s2.u0.pw1','tower.s2.u4.dw','tower.s3.u0.scdw','tower.s3.u3.dw','tower.head']:
    print(f'{k:22s} rel_l2={C.rel_l2(e16.tensor(k).float(), e32.tensor(k)):.3e}')
print('x512 rel_l2', C.rel_l2(e16.x512, e32.x512), 'loss', s16[0].item(), s32[0].item())
g16, g32 = e16.dyn.to_dict(e16.g_dyn), e32.dyn.to_dict(e32.g_dyn)
rows = sorted((C.rel_l2(g16[k], g32[k]), k) for k in g32 if g32[k].abs().max() > 1e-9)
print('grad rel_l2 median', rows[len(rows)//2], 'p90', rows[int(len(rows)*.9)], 'max', rows[-1])
print('cos(g16,g32) whole arena', torch.nn.functional.cosine_similarity(e16.g_dyn, e32.g_dyn, dim=0).item())
