"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the oracle on the same
seeded inputs (SURVEY §8c,d).  fp32 "parity mode" is compared with the fp64 oracle; bf16 "perf mode"
against the same oracle with bf16-level tolerances; full-size runs are checked through
size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import model, ppo, spec
from tests import common as C

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason='needs a CUDA device')]

H, W = 90, 120


def _engine(B, dtype, h=H, w=W):
    from cdra.engine import Engine
    return Engine(B, h, w, dtype=dtype, image_u8=True, device='cuda')


def _dev(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.fixture(scope='module')
def params():
    return C.fresh_params(torch.float64)


TAPS = ('tower.stem', 'tower.pool', 'tower.s1.u0.pw1', 'tower.s1.u0.dw', 'tower.s1.u0.scdw', 'tower.s1.u1.pw1', 'tower.s1.u3.dw',
        'tower.s2.u0.pw1', 'tower.s2.u4.dw', 'tower.s3.u0.scdw', 'tower.s3.u3.dw', 'tower.head')


def test_forward_fp32_matches_oracle(built_libs, params):
    B = 8
    dyn, pol, val = params
    eng = _engine(B, 'f32')
    C.load_engine(eng, dyn, pol, val)
    obs = C.synthetic_obs(B, H, W, seed=1234)
    out = eng.dynamics_forward(_dev(obs)).clone()
    torch.cuda.synchronize()
    taps, bs = {}, model.BNState()
    ref = model.dynamics_forward(dyn, C.oracle_obs(obs), True, bs, taps)
    for k in TAPS:
        assert C.rel_max(eng.tensor(k)[:B], taps[k]) < 1e-4, k             # fp32 tolerance: 1e-4 of the tensor's max
    assert C.rel_max(eng.tensor('tower.gap'), taps['tower.gap']) < 1e-4
    assert C.rel_max(out, ref) < 2e-4
    new, got = bs.apply_moving(dyn), eng.dyn_state.to_dict()
    assert max(C.rel_max(got[k], new[k]) for k in got) < 1e-5
    # index ops are bit-exact: the shortcut half of a stride-1 unit is a permutation of the input's left half
    x, y = eng.tensor('tower.s2.u0.out'), eng.tensor('tower.s2.u1.out')
    c = x.shape[-1]
    for j, q in enumerate(model.shuffle_perm(c)):
        if q < c // 2:
            assert torch.equal(y[..., j], x[..., q])


def test_forward_fp32_trained_weights(built_libs):
    """same check on the reference's shipped stage-s5-curriculum agent (realistic BN statistics / saturation)"""
    B = 4
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B, 'f32')
    C.load_engine(eng, dyn, pol, val)
    obs = C.synthetic_obs(B, H, W, seed=77)
    out = eng.dynamics_forward(_dev(obs)).clone()
    ref = model.dynamics_forward(dyn, C.oracle_obs(obs), True)
    assert C.rel_max(out, ref) < 1e-3


def _check_grads(eng, arena, flat, ref_grads, head):
    rows = C.grad_report(arena, flat, ref_grads)
    if head:
        assert max(r[2] for r in rows) < 1e-3, sorted(rows, key=lambda r: -r[2])[:3]
        return
    tail = [r for r in rows if not r[0].startswith('tower.')]
    assert max(r[2] for r in tail) < 2e-3, sorted(tail, key=lambda r: -r[2])[:3]
    # tower: ReLU6 / max-pool decisions that sit within rounding distance of a boundary may flip between
    # the fp32 kernels and the fp64 oracle (any two implementations differ there); relative L2 is robust
    l2 = sorted(r[1] for r in rows)
    assert l2[len(l2) // 2] < 1e-2 and l2[int(len(l2) * 0.9)] < 3e-2 and l2[-1] < 0.2, (l2[len(l2) // 2], l2[-1])
    # the layers closest to the loss see (almost) no flipped decision upstream: tight agreement there.  One borderline
    # ReLU6 decision of the head conv itself moves these gradients by ~1/(B*12) = 1 %, and the order of the fp32 / fp64
    # atomics decides it run by run, hence 3e-2 rather than 1e-2
    late = [r for r in rows if r[0].startswith(('tower.head', 'tower.s3.u3.pw2', 'tower.s3.u3.dw'))]
    assert max(r[1] for r in late) < 3e-2, late


def test_policy_pass_fp32(built_libs, params):
    B = 8
    dyn, pol, val = params
    eng = _engine(B, 'f32')
    C.load_engine(eng, dyn, pol, val)
    obs, bt = C.synthetic_obs(B, H, W, seed=21), C.synthetic_batch(B, seed=22)
    sc = C.policy_step_engine(eng, _dev(obs), _dev(bt)).cpu()
    ref = C.policy_step_oracle(dyn, pol, obs, bt)
    # north-star tolerance: PPO loss within rtol 1e-4 of the reference arithmetic
    assert abs(sc[0].item() - ref['loss'].item()) <= 1e-4 * abs(ref['loss'].item()) + 1e-6
    names = ['loss_total', 'loss_policy', 'loss_entropy', 'loss_speed_policy', 'loss_similarity_policy', 'ratio', 'log_prob',
             'entropy', 'speed_pi', 'similarity_pi']
    for i, n in enumerate(names):
        assert abs(sc[i].item() - ref['scalars'][n].item()) <= 1e-4 * abs(ref['scalars'][n].item()) + 2e-6, n
    _check_grads(eng, eng.pol, eng.g_pol, ref['g_head'], True)
    _check_grads(eng, eng.dyn, eng.g_dyn, ref['g_dyn'], False)


def test_value_pass_fp32(built_libs, params):
    B = 8
    dyn, pol, val = params
    eng = _engine(B, 'f32')
    C.load_engine(eng, dyn, pol, val)
    obs, bt = C.synthetic_obs(B, H, W, seed=31), C.synthetic_batch(B, seed=32)
    sc = C.value_step_engine(eng, _dev(obs), _dev(bt)).cpu()
    ref = C.value_step_oracle(dyn, val, obs, bt)
    assert abs(sc[0].item() - ref['loss'].item()) <= 1e-4 * abs(ref['loss'].item()) + 1e-6
    _check_grads(eng, eng.val, eng.g_val, ref['g_head'], True)
    _check_grads(eng, eng.dyn, eng.g_dyn, ref['g_dyn'], False)


def test_bf16_mode_tracks_oracle(built_libs, params):
    """perf mode: bf16 activation storage, fp32 accumulate.  Each stored tensor carries ~2^-9 relative rounding noise.
    With the reference's trained weights that noise stays at the 1-2 % level through all 50 layers (measured:
    profiles/r1_bf16_vs_f32.txt); a randomly initialised BatchNorm tower is chaotic (perturbations grow ~1.3x per unit),
    so there only the early layers and the loss are held to a tolerance."""
    B = 8
    obs, bt = C.synthetic_obs(B, H, W, seed=41), C.synthetic_batch(B, seed=42)
    # (a) trained agent: every tap within 5 % (relative L2) of the fp64 oracle
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B, 'bf16')
    C.load_engine(eng, dyn, pol, val)
    out = eng.dynamics_forward(_dev(obs)).clone()
    taps = {}
    ref = model.dynamics_forward(dyn, C.oracle_obs(obs), True, model.BNState(), taps)
    for k in TAPS:
        assert C.rel_l2(eng.tensor(k)[:B].float(), taps[k]) < 5e-2, k
    assert C.rel_l2(out, ref) < 5e-2
    # (b) random init: early layers at bf16 precision, loss in the right place, finite gradients
    dyn, pol, val = params
    C.load_engine(eng, dyn, pol, val)
    sc = C.policy_step_engine(eng, _dev(obs), _dev(bt)).cpu()
    ref = C.policy_step_oracle(dyn, pol, obs, bt)
    taps = {}
    model.dynamics_forward(dyn, C.oracle_obs(obs), True, model.BNState(), taps)
    for k in ('tower.stem', 'tower.pool', 'tower.s1.u0.pw1', 'tower.s1.u0.dw', 'tower.s1.u0.scdw'):
        assert C.rel_l2(eng.tensor(k)[:B].float(), taps[k]) < 2e-2, k
    assert abs(sc[0].item() - ref['loss'].item()) < 0.25 * max(1.0, abs(ref['loss'].item()))
    assert torch.isfinite(eng.g_dyn).all() and torch.isfinite(eng.g_pol).all()


def test_gae_full_size_bit_exact(built_libs):
    """BASELINE config sizes: bs 512 x T 256 trajectories; exponents and bases bit-exact vs scipy.lfilter path"""
    eng = _engine(2, 'f32', 42, 58)
    rng = np.random.RandomState(0)
    bs, T = 512, 256
    rew = (rng.randn(bs, T) * 2 + 1).clip(-10, 30).astype('f')
    vbe = np.stack([rng.rand(bs, T) * 2 - 1, rng.rand(bs, T) * 6], -1).astype('f')
    last = np.stack([rng.rand(bs) * 2 - 1, rng.rand(bs) * 6], -1).astype('f')
    last[::7] = 0
    rb, adv = eng.gae(torch.tensor(rew).cuda(), torch.tensor(vbe).cuda(), torch.tensor(last).cuda(), 0.9999, 0.999, 2.0)
    rb, adv = rb.cpu().numpy(), adv.cpu().numpy()
    for i in range(0, bs, 3):
        r_ref, a_ref, _ = ppo.end_trajectory(rew[i], vbe[i], last[i], 0.9999, 0.999, 2.0)
        assert np.array_equal(rb[i][:, 1], r_ref[:, 1])                     # integer exponents: exact
        assert np.abs(rb[i][:, 0] - r_ref[:, 0]).max() <= 1.2e-7           # bases: <= 1 ulp (device pow)
        assert np.abs(adv[i] - a_ref).max() < 1e-6
    # long-horizon stress (config 5): T = 512
    rb2, adv2 = eng.gae(torch.tensor(np.tile(rew[:8], (1, 2))).cuda(), torch.tensor(np.tile(vbe[:8], (1, 2, 1))).cuda(),
                        torch.tensor(last[:8]).cuda(), 0.9999, 0.999, 2.0)
    r_ref, a_ref, _ = ppo.end_trajectory(np.tile(rew[3], 2), np.tile(vbe[3], (2, 1)), last[3], 0.9999, 0.999, 2.0)
    assert np.array_equal(rb2[3].cpu().numpy()[:, 1], r_ref[:, 1])
    assert np.abs(adv2[3].cpu().numpy() - a_ref).max() < 1e-6


def test_clip_adam_and_gather(built_libs):
    eng = _engine(2, 'f32', 42, 58)
    arena = eng.pol
    g = torch.Generator().manual_seed(3)
    p0 = {n: torch.randn(s, generator=g) for n, s in zip(arena.names, arena.shapes)}
    g0 = {n: torch.randn(s, generator=g) * (3.0 if i % 2 else 0.01) for i, (n, s) in enumerate(zip(arena.names, arena.shapes))}
    arena.load_dict(p0)
    for n in arena.names:
        arena.view(n, eng.g_pol).copy_(g0[n])
    pr = {k: v.clone().double() for k, v in p0.items()}
    m = {k: torch.zeros_like(v) for k, v in pr.items()}
    v = {k: torch.zeros_like(t) for k, t in pr.items()}
    for step in (1, 2, 3):
        eng.clip_adam('pol', 3e-4, clip_norm=1.0)
        ppo.apply_step(pr, {k: t.double() for k, t in g0.items()}, m, v, step, 3e-4, clip=1.0)
    got = arena.to_dict()
    assert max((got[k].double().cpu() - pr[k]).abs().max().item() for k in got) < 2e-6
    # dynamics optimiser: no clipping (core/carla_agent.py:386-388)
    eng.g_dyn.normal_(generator=None)
    before = eng.dyn.flat.clone()
    eng.clip_adam('dyn', 1e-3)
    step = (eng.dyn.flat - before).abs()
    assert abs(step.max().item() - 1e-3) < 1e-5           # first Adam step has magnitude lr for every coordinate
    src = torch.randint(0, 256, (64, 4 * 90 * 120 * 3), dtype=torch.uint8, device='cuda')
    idx = torch.randperm(64, device='cuda')[:32]
    out = torch.empty(32, src.shape[1], dtype=torch.uint8, device='cuda')
    eng.gather_rows(src, idx, out)
    assert torch.equal(out, src[idx])
    srcs = [src, torch.randn(64, 4, 9, device='cuda'), torch.randn(64, 1, device='cuda'), torch.randn(64, 2, device='cuda')]
    outs = [torch.empty((32,) + tuple(t.shape[1:]), dtype=t.dtype, device='cuda') for t in srcs]
    eng.gather_rows_multi(srcs, idx, outs)                # the whole minibatch in one launch
    for t, o in zip(srcs, outs):
        assert torch.equal(o, t[idx])


def test_full_size_properties_bf16(built_libs, params):
    """BASELINE config 2 minibatch (B = 512, 90x120, bf16): size-independent properties"""
    B = 512
    dyn, pol, val = params
    eng = _engine(B, 'bf16')
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, H, W, seed=51)), _dev(C.synthetic_batch(B, seed=52))
    sc = C.policy_step_engine(eng, obs, bt)
    assert torch.isfinite(sc[:10]).all() and torch.isfinite(eng.g_dyn).all() and torch.isfinite(eng.g_pol).all()
    # (1) every BatchNorm output is normalised per (slice, channel): mean 0 / variance 1 before gamma, beta
    raw = eng.tensor('tower.s2.u3.pw1').float().view(4, B * 6 * 8, -1)
    mu, var = raw.mean(1), raw.var(1, unbiased=False)
    g_, b_ = eng.dyn.view('tower.s2.u3.pw1.g'), eng.dyn.view('tower.s2.u3.pw1.be')
    z = (raw - mu[:, None]) * torch.rsqrt(var[:, None] + 1e-3)
    assert z.mean(1).abs().max() < 1e-3 and (z.var(1, unbiased=False) - var / (var + 1e-3)).abs().max() < 1e-2
    # (2) channel shuffle / pass-through is bit-exact at full size
    x, y = eng.tensor('tower.s3.u1.out'), eng.tensor('tower.s3.u2.out')
    c = x.shape[-1]
    perm = model.shuffle_perm(c)
    for j in range(0, c, 7):
        if perm[j] < c // 2:
            assert torch.equal(y[..., j], x[..., perm[j]])
    # (3) gradients of parameters that only shift a BatchNorm input are (numerically) zero
    gb = eng.dyn.view('tower.s1.u1.pw1.b', eng.g_dyn).abs().max().item()
    gw = eng.dyn.view('tower.s1.u1.pw1.w', eng.g_dyn).abs().max().item()
    assert gb < 5e-2 * gw
    # (4) backward is linear in d_out: scaling the upstream gradient scales every parameter gradient
    g1 = eng.g_dyn.clone()
    eng.dynamics_backward(obs, eng.d_x512 * 2.0)
    assert C.rel_l2(eng.g_dyn, 2.0 * g1) < 2e-2
    # (5) a few SGD steps on a fixed minibatch reduce the value loss
    losses = []
    for _ in range(4):
        s = C.value_step_engine(eng, obs, bt)
        losses.append(s[0].item())
        eng.clip_adam('dyn', 3e-4); eng.clip_adam('val', 3e-4, clip_norm=1.0)
    assert losses[-1] < losses[0]


def test_config4_value_pass_permutation_bf16(built_libs):
    """The same permutation / repeatability property for the VALUE pass at the 180x240 geometry of BASELINE config 4
    (B = 256 per launch): the row-banded depthwise / stem kernels and the 4x larger row counts of every GEMM launch."""
    B, h, w = 256, 180, 240
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B, 'bf16', h, w)
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, h, w, seed=71)), _dev(C.synthetic_batch(B, seed=72))
    l0 = C.value_step_engine(eng, obs, bt)[0].item()
    gd0, gv0 = eng.g_dyn.clone(), eng.g_val.clone()
    l1 = C.value_step_engine(eng, obs, bt)[0].item()
    rep_d, rep_v = C.rel_l2(eng.g_dyn, gd0), C.rel_l2(eng.g_val, gv0)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(9)).cuda()
    l2 = C.value_step_engine(eng, {k: v[perm].contiguous() for k, v in obs.items()}, {k: v[perm].contiguous() for k, v in bt.items()})[0].item()
    prm_d, prm_v = C.rel_l2(eng.g_dyn, gd0), C.rel_l2(eng.g_val, gv0)
    print(f'repeat: loss {l0:.7f} / {l1:.7f}, g_dyn {rep_d:.2e}, g_val {rep_v:.2e};  permuted: loss {l2:.7f}, g_dyn {prm_d:.2e}, g_val {prm_v:.2e}')
    assert torch.isfinite(eng.g_dyn).all() and eng.g_dyn.abs().max() > 0
    assert abs(l1 - l0) <= 1e-5 * abs(l0) and rep_d < 1e-3 and rep_v < 1e-4
    assert abs(l2 - l0) <= 1e-5 * abs(l0) and prm_d < 1e-3 and prm_v < 1e-4


def test_full_size_permutation_and_repeatability_bf16(built_libs):
    """BASELINE config 2 minibatch (B = 512, 90x120, bf16), whole policy pass on the trained stage-s5 weights (a freshly
    initialised tower amplifies a single bf16 rounding flip ~30x per unit -- profiles/dbg_repeat.py -- so only a trained
    network has a meaningful end-to-end tolerance): (a) the same step twice gives the same loss and
    gradients up to the order of the fp32 / fp64 atomics; (b) permuting the samples of the minibatch (observations and PPO
    batch alike) changes neither: BatchNorm statistics, the loss and every parameter gradient are sums over the batch.  The
    permutation moves every sample into another 128-row tile, another CTA and another ring slot of every kernel on the path,
    so a tile-, band- or slot-dependent error shows up here at the size the bench runs."""
    B = 512
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B, 'bf16')
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, H, W, seed=61)), _dev(C.synthetic_batch(B, seed=62))
    l0 = C.policy_step_engine(eng, obs, bt)[0].item()
    gd0, gp0 = eng.g_dyn.clone(), eng.g_pol.clone()
    l1 = C.policy_step_engine(eng, obs, bt)[0].item()
    rep_d, rep_p = C.rel_l2(eng.g_dyn, gd0), C.rel_l2(eng.g_pol, gp0)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(7)).cuda()
    obs_p = {k: v[perm].contiguous() for k, v in obs.items()}
    bt_p = {k: v[perm].contiguous() for k, v in bt.items()}
    l2 = C.policy_step_engine(eng, obs_p, bt_p)[0].item()
    prm_d, prm_p = C.rel_l2(eng.g_dyn, gd0), C.rel_l2(eng.g_pol, gp0)
    print(f'repeat: loss {l0:.7f} / {l1:.7f}, g_dyn {rep_d:.2e}, g_pol {rep_p:.2e};  permuted: loss {l2:.7f}, g_dyn {prm_d:.2e}, g_pol {prm_p:.2e}')
    # measured on B200: repeat g_dyn 2.3e-5 / g_pol 2.6e-7, permuted 3.5e-5 / 2.7e-7, loss equal to 1e-7 (order of the fp32 / fp64
    # atomics in the sums and weight gradients); the bounds leave ~30x
    assert abs(l1 - l0) <= 1e-5 * abs(l0) and rep_d < 1e-3 and rep_p < 1e-4
    assert abs(l2 - l0) <= 1e-5 * abs(l0) and prm_d < 1e-3 and prm_p < 1e-4

def test_stem_backward_tensor_core_vs_cuda_core(built_libs, params):
    """the one-pass tensor-core stem backward (max-pool backward + BN backward folded into the weight gradient by
    linearity, v2_stem.cuh) against the CUDA-core kernels that materialise every intermediate, on identical inputs"""
    B = 8
    dyn, pol, val = params
    eng = _engine(B, 'bf16')
    C.load_engine(eng, dyn, pol, val)
    obs, bt = _dev(C.synthetic_obs(B, H, W, seed=61)), _dev(C.synthetic_batch(B, seed=62))
    C.policy_step_engine(eng, obs, bt)
    g_step = eng.dyn.to_dict(eng.g_dyn.clone())
    g_new = eng.dyn.to_dict(eng.debug_stem_backward(obs, False))
    g_old = eng.dyn.to_dict(eng.debug_stem_backward(obs, True))
    for k in ('tower.stem.w', 'tower.stem.g', 'tower.stem.be'):
        assert C.rel_l2(g_new[k], g_step[k]) < 5e-3, k      # replay == the step (up to the order of the bf16 / fp64 atomics)
        assert C.rel_l2(g_new[k], g_old[k]) < 2e-2, (k, C.rel_l2(g_new[k], g_old[k]))           # bf16 storage of d stem in the legacy path
    assert g_new['tower.stem.b'].abs().max().item() == 0.0                                      # bias before a training-mode BN


@pytest.mark.parametrize('tensor_core', [0, 1])
def test_dense_gemm_matches_torch(built_libs, tensor_core):
    """the GEMM under the GRUs / trunk / control branches, all four operand orientations, ragged sizes, unaligned bases"""
    from cdra import _lib
    lib = _lib.load()
    g = torch.Generator(device='cuda').manual_seed(5)
    tol = 2e-3 if tensor_core else 1e-5          # TF32 operands (10-bit mantissa) vs fp32
    for (M, N, K) in ((2048, 768, 768), (512, 512, 352), (8, 96, 16), (70, 130, 33), (352, 512, 512)):
        for ta in (0, 1):
            for tb in (0, 1):
                for off in (0, 1):               # off = 1: operands start at an odd float (no 16-byte alignment)
                    A = torch.randn((K, M) if ta else (M, K), generator=g, device='cuda')
                    Bm = torch.randn((N, K) if tb else (K, N), generator=g, device='cuda')
                    bias = torch.randn(N, generator=g, device='cuda')
                    bufA, bufB = torch.zeros(A.numel() + 1, device='cuda'), torch.zeros(Bm.numel() + 1, device='cuda')
                    bufA[off:off + A.numel()] = A.flatten(); bufB[off:off + Bm.numel()] = Bm.flatten()
                    Cm = torch.ones(M, N, device='cuda')
                    ref = (A.double().t() if ta else A.double()) @ (Bm.double().t() if tb else Bm.double()) + bias.double() + 1.0
                    pa, pb = bufA.data_ptr() + 4 * off, bufB.data_ptr() + 4 * off
                    rc = lib.cdra_debug_gemm(ta, tb, pa, A.shape[1], pb, Bm.shape[1], _lib.ptr(Cm), N, _lib.ptr(bias), M, N, K, 1, tensor_core, None)
                    assert rc == 0
                    torch.cuda.synchronize()
                    err = ((Cm.double() - ref).abs().max() / ref.abs().max()).item()
                    assert err < tol, (M, N, K, ta, tb, off, err)


@pytest.mark.parametrize('shape', [(128, 128, 64), (256, 128, 128), (128, 256, 240), (192, 256, 48)])
def test_tcgen05_selftest(built_libs, shape):
    """tcgen05.mma with MN-major 128B-swizzled operands and TMEM accumulators (the weight-gradient product X^T Y)"""
    from cdra import _lib
    lib = _lib.load()
    rows, Mw, Nw = shape
    g = torch.Generator(device='cuda').manual_seed(rows + Nw)
    X = torch.randn(rows, Mw, generator=g, device='cuda').bfloat16()
    Y = torch.randn(rows, Nw, generator=g, device='cuda').bfloat16()
    Cm = torch.full((Mw, Nw), float('nan'), device='cuda')
    assert lib.cdra_debug_umma_selftest(_lib.ptr(X), _lib.ptr(Y), _lib.ptr(Cm), rows, Mw, Nw, None) == 0
    torch.cuda.synchronize()
    ref = X.double().t() @ Y.double()
    assert ((Cm.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5


def test_high_res_tower_bf16(built_libs, params):
    """BASELINE config 4 geometry (180x240): stage-1 frames exceed shared memory, so the depthwise kernels work on row BANDS
    of a frame (halo rows re-read, out-of-frame tile rows re-zeroed per band); perf mode must track the fp32 parity mode"""
    B, h, w = 2, 180, 240
    dyn, pol, val = C.trained_params(torch.float64)
    obs, bt = _dev(C.synthetic_obs(B, h, w, seed=81)), _dev(C.synthetic_batch(B, seed=82))
    outs = {}
    for dt in ('f32', 'bf16'):
        eng = _engine(B, dt, h, w)
        C.load_engine(eng, dyn, pol, val)
        sc = C.policy_step_engine(eng, obs, bt)
        assert torch.isfinite(sc[:10]).all() and torch.isfinite(eng.g_dyn).all()
        outs[dt] = (eng.x512.clone(), sc[0].item())
    assert C.rel_l2(outs['bf16'][0], outs['f32'][0]) < 5e-2
    assert abs(outs['bf16'][1] - outs['f32'][1]) < 0.05 * max(1.0, abs(outs['f32'][1]))


@pytest.mark.parametrize('shape', [(128, 64, 64), (128, 128, 128), (256, 240, 128), (128, 256, 192)])
def test_tcgen05_selftest_k_major(built_libs, shape):
    """tcgen05.mma with K-major 128B-swizzled operands (the forward / data-gradient product A B^T)"""
    from cdra import _lib
    lib = _lib.load()
    Mw, Nw, Kw = shape
    g = torch.Generator(device='cuda').manual_seed(Mw + Nw + Kw)
    A = torch.randn(Mw, Kw, generator=g, device='cuda').bfloat16()
    Bm = torch.randn(Nw, Kw, generator=g, device='cuda').bfloat16()
    Cm = torch.full((Mw, Nw), float('nan'), device='cuda')
    assert lib.cdra_debug_umma_selftest_k(_lib.ptr(A), _lib.ptr(Bm), _lib.ptr(Cm), Mw, Nw, Kw, None) == 0
    torch.cuda.synchronize()
    ref = A.double() @ Bm.double().t()
    assert ((Cm.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5


def test_pointwise_forward_tcgen05_vs_mma(built_libs):
    """the tcgen05 forward kernel of the plain-output pointwise layers (K-major swizzled operands, TMEM accumulator)
    against the mma.sync kernel on identical inputs"""
    from cdra import _lib
    lib = _lib.load()
    B = 8
    dyn, pol, val = C.trained_params(torch.float64)
    eng = _engine(B, 'bf16')
    C.load_engine(eng, dyn, pol, val)
    obs = _dev(C.synthetic_obs(B, H, W, seed=91))
    names = ('tower.s1.u0.pw1', 'tower.s1.u2.pw1', 'tower.s2.u3.pw1', 'tower.s3.u1.pw1', 'tower.head')
    res = {}
    try:
        for v in (0, 1):
            assert lib.cdra_debug_set(b'fwd_tc', v) == 0
            out = eng.dynamics_forward(obs).clone()
            torch.cuda.synchronize()
            res[v] = ({n: eng.tensor(n).float().clone() for n in names}, out)
    finally:
        lib.cdra_debug_set(b'fwd_tc', -1)
    # first layer: same bf16 operands, only the fp32 accumulation order differs (a few results round the other way)
    assert C.rel_l2(res[1][0]['tower.s1.u0.pw1'], res[0][0]['tower.s1.u0.pw1']) < 5e-4
    for n in names[1:]:
        assert C.rel_l2(res[1][0][n], res[0][0][n]) < 3e-2, n        # bf16 rounding differences carried through the layers
    assert C.rel_l2(res[1][1], res[0][1]) < 1e-2
