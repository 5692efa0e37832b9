"""Run-to-run repeatability of the bf16 path at the benchmark size: which tensors differ between two identical calls."""
import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/carla-driving-rl-agent_b200')
from cdra.engine import Engine
from tests import common as C
B = int(os.environ.get('B', 512))
eng = Engine(B, 90, 120, dtype='bf16', image_u8=True, device='cuda')
dyn, pol, val = C.trained_params(torch.float64) if os.environ.get('TRAINED') else C.fresh_params(torch.float64)
C.load_engine(eng, dyn, pol, val)
obs = {k: v.cuda() for k, v in C.synthetic_obs(B, 90, 120, seed=61).items()}
bt = {k: v.cuda() for k, v in C.synthetic_batch(B, seed=62).items()}
taps = ['tower.stem', 'tower.pool', 'tower.s1.u0.pw1', 'tower.s1.u0.dw', 'tower.s1.u0.out', 'tower.s1.u1.pw1', 'tower.s1.u1.dw', 'tower.s1.u1.out', 'tower.s1.u3.out',
        'tower.s2.u0.out', 'tower.s2.u7.out', 'tower.s3.u0.out', 'tower.s3.u3.out', 'tower.head']
def run():
    x = eng.dynamics_forward(obs).clone()
    t = {n: eng.tensor(n).clone() for n in taps}
    return x, t
x0, t0 = run(); x1, t1 = run()
print('x512 max abs diff', (x0 - x1).abs().max().item(), 'rel-L2', C.rel_l2(x0, x1))
for n in taps:
    d = (t0[n].float() - t1[n].float())
    print(f'{n:20s} differing elements {int((d != 0).sum())} of {d.numel()}  max abs {d.abs().max().item():.3e}')
s0 = C.policy_step_engine(eng, obs, bt); g0 = eng.g_dyn.clone(); p0 = eng.g_pol.clone()
s1 = C.policy_step_engine(eng, obs, bt)
print('loss', s0[0].item(), s1[0].item(), 'g_dyn rel-L2', C.rel_l2(eng.g_dyn, g0), 'g_pol', C.rel_l2(eng.g_pol, p0))
worst = []
for n, v in eng.dyn.items(eng.g_dyn):
    a = eng.dyn.view(n, g0)
    worst.append((C.rel_l2(v, a), n, a.norm().item()))
worst.sort(reverse=True)
for w in worst[:12]: print('  %.3e %-28s |g| %.3e' % w)
