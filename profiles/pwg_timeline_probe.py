"""Role timeline of one launch of the tcgen05 forward GEMM family (head conv) and of the fused pointwise backward (s1.u0.pw1)
at the benchmark size: python profiles/pwg_timeline_probe.py   (needs a B200; CDRA_TIMELINE=1 is set here)."""
import os, sys, ctypes, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/carla-driving-rl-agent_b200')
os.environ['CDRA_TIMELINE'] = '1'
from cdra.engine import Engine
from tests import common as C
for B in (512,):
    eng = Engine(B, 90, 120, dtype='bf16', image_u8=True, device='cuda')
    dyn, pol, val = C.fresh_params(torch.float64)
    C.load_engine(eng, dyn, pol, val)
    obs = {k: v.cuda() for k, v in C.synthetic_obs(B, 90, 120, seed=1).items()}
    bt = {k: v.cuda() for k, v in C.synthetic_batch(B, seed=2).items()}
    for _ in range(3): C.policy_step_engine(eng, obs, bt)
    torch.cuda.synchronize()
    ts = (ctypes.c_ulonglong * 64)()
    eng.lib.cdra_debug_timeline(ts)
    t = list(ts)
    names = ['start', 'prologue done', 'pdl_wait done', 'weights landed (MMA thr)', 'first full (MMA thr)', 'first tm_full (epi)', 'first item copied', 'last item done', 'flush done', 'final sync', 'last_cta elected', 'finalize done']
    print('B =', B, '(head conv = last pwg_fwd launch, block (0,0))')
    for i, n in enumerate(names):
        print(f'  {n:28s} {(t[i] - t[0]) / 1e3:8.2f} us')
    print('  pwg_dgrad (last launch = first stage-3 layer of the backward... see note), tile 5 epilogue: loop top %.2f, accumulator ready %.2f, staged %.2f, copied out %.2f (us, same clock as above)' % tuple((t[i] - t[0]) / 1e3 for i in (12, 13, 14, 15)))
    names3 = ['start', 'prologue done', 'pdl_wait done', 'first full (transform)', 'first stg_full (MMA)', 'first tm_full (epi)', 'tile 20: epilogue channel loop done', 'last tile done (epi)', 'roles joined', 'dW flush done', 'MMA tile 8 staged', 'BN param grads done']
    print('fused backward, last launch (s1.u0.pw1), block 0')
    for i, n in enumerate(names3):
        print(f'  {n:28s} {(t[16 + i] - t[16]) / 1e3:8.2f} us')
    # steady state of the same launch: tile 20 of block 0, role by role (us since the launch's start)
    u = lambda i: (t[48 + i - 16] - t[16]) / 1e3 if i >= 16 else (t[16 + i] - t[16]) / 1e3
    print('  tile 20  TMA producer : loop top %.2f, ring slot free %.2f' % (u(16), u(17)))
    print('  tile 20  MMA issuer   : loop top %.2f, operands staged %.2f, accumulator free %.2f, both GEMMs issued %.2f' % (u(18), u(19), u(20), u(21)))
    print('  tile 20  transform    : loop top %.2f, rows landed %.2f, staging free %.2f, handed over %.2f; tile 21 handed over %.2f' % (u(22), u(23), u(24), u(25), u(26)))
    print('  tile 20  epilogue     : loop top %.2f, staging rows free %.2f, accumulator ready %.2f, channels done %.2f, stores issued %.2f; tile 22 (same group) stores issued %.2f' % (u(27), u(28), u(29), u(30), u(31), u(12)))
    names4 = ['start', 'prologue done', 'pdl_wait done', 'tile 2: loop top', 'tile 2: raw rows landed', 'tile 2: transformed + CTA barrier', 'tile 2: MMA complete', 'tile 2: TMEM -> staging + CTA barrier', 'tile 2: stored + statistics', 'all tiles done', 'statistics flushed']
    print('pw_fwd_tc, last launch of the backward-free forward (s3.u0.pw1 is pwg; this is the last stage-2 pw1), block 0')
    for i, n in enumerate(names4):
        print(f'  {n:40s} {(t[32 + i] - t[32]) / 1e3:8.2f} us')
