"""Debug aid: tcgen05 forward kernel (plain-output pointwise layers) vs the mma.sync kernel, same inputs, one process."""
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
obs = {k: v.cuda() for k, v in C.synthetic_obs(B, H, W, seed=41).items()}
names = ['tower.s1.u0.pw1', 'tower.s1.u2.pw1', 'tower.s2.u0.pw1', 'tower.s2.u3.pw1', 'tower.s2.u7.pw1', 'tower.s3.u1.pw1', 'tower.head']
res = {}
for v in (0, 1):
    assert lib.cdra_debug_set(b'fwd_tc', v) == 0
    out = eng.dynamics_forward(obs).clone(); torch.cuda.synchronize()
    res[v] = ({n: eng.tensor(n).float().clone() for n in names}, out, {k: t.clone() for k, t in eng.dyn_state.to_dict().items()})
lib.cdra_debug_set(b'fwd_tc', -1)
for n in names:
    print(f'{n:20s} tc-vs-mma rel_l2 {C.rel_l2(res[1][0][n], res[0][0][n]):.3e}  finite {bool(torch.isfinite(res[1][0][n]).all())}')
print('out512 rel_l2', C.rel_l2(res[1][1], res[0][1]))
print('moving stats worst rel_max', max(C.rel_max(res[1][2][k], res[0][2][k]) for k in res[0][2]))
