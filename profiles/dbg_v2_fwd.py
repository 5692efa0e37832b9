"""Debug aid: v2 (bf16) tower forward vs the fp64 oracle, tap by tap (relative L2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'carla-driving-rl-agent_b200')]
import torch
from oracle import model, spec
from tests import common as C
from cdra.engine import Engine

B, H, W = 8, 90, 120
obs = C.synthetic_obs(B, H, W, seed=41)
dev = {k: v.cuda() for k, v in obs.items()}
for label, params in (('trained', C.trained_params(torch.float64)), ('random', C.fresh_params(torch.float64))):
    dyn, pol, val = params
    eng = Engine(B, H, W, dtype='bf16', image_u8=True, device='cuda')
    C.load_engine(eng, dyn, pol, val)
    out = eng.dynamics_forward(dev).clone()
    torch.cuda.synchronize()
    taps = {}
    ref = model.dynamics_forward(dyn, C.oracle_obs(obs), True, model.BNState(), taps)
    print('==', label)
    names = ['tower.pool']
    for name, stride, cin, c in spec.tower_units():
        names += [name + '.pw1', name + '.dw'] + ([name + '.scdw'] if stride == 2 else [])
    names += ['tower.head']
    worst = 0
    for k in names:
        e = C.rel_l2(eng.tensor(k)[:B].float(), taps[k])
        worst = max(worst, e)
        print(f'{k:24s} {e:.4f}')
    print('gap', C.rel_l2(eng.tensor('tower.gap'), taps['tower.gap']), 'out512', C.rel_l2(out, ref), 'worst', worst)
    new, got = model.BNState(), None
