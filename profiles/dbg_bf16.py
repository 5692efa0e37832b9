import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'carla-driving-rl-agent_b200')
from cdra.engine import Engine
from oracle import model
from tests import common as C
B, H, W = 8, 90, 120
dyn, pol, val = C.fresh_params(torch.float64)
eng = Engine(B, H, W, dtype='bf16', image_u8=True, device='cuda')
C.load_engine(eng, dyn, pol, val)
obs, bt = C.synthetic_obs(B, H, W, seed=41), C.synthetic_batch(B, seed=42)
dev = lambda d: {k: v.cuda() for k, v in d.items()}
sc = C.policy_step_engine(eng, dev(obs), dev(bt)).cpu()
ref = C.policy_step_oracle(dyn, pol, obs, bt)
taps = {}
model.dynamics_forward(dyn, C.oracle_obs(obs), True, model.BNState(), taps)
for k in ['tower.stem','tower.pool','tower.s1.u0.pw1','tower.s1.u0.dw','tower.s1.u0.scdw','tower.s1.u1.pw1','tower.s1.u3.dw','tower.s2.u0.pw1','tower.s2.u4.dw','tower.s3.u0.scdw','tower.s3.u3.dw','tower.head']:
    print(f'{k:22s} rel_l2={C.rel_l2(eng.tensor(k)[:B].float(), taps[k]):.3e}')
print('x512 rel_l2', C.rel_l2(eng.x512, ref['x512']), 'loss', sc[0].item(), ref['loss'].item())
rows = C.grad_report(eng.dyn, eng.g_dyn, ref['g_dyn'])
l2 = sorted(r[1] for r in rows); print('grad l2 median', l2[len(l2)//2], 'p90', l2[int(len(l2)*.9)], 'max', l2[-1])
for r in sorted(rows, key=lambda r: -r[1])[:12]: print(r)
for r in rows:
    if r[0].startswith(('tower.head','tower.s3.u3','gru.image','trunk')): print(r)
