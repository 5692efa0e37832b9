#!/usr/bin/env python
"""PPO update-step throughput of the B200-native hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (libcdra, sm_100a), BASELINE config C2 per GPU
    python bench.py --config C4 --gpus 2 ...                 # the other BASELINE.json configurations by name (C2 | C4 | C5)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle restatement of the
                                                             # reference's TF/Keras path on the host cores

A "step" is one SGD minibatch index of the PPO update over the global minibatch (bs samples per GPU):
policy pass (shared-trunk forward+backward, policy head + clipped-surrogate/entropy/aux loss, gradient
all-reduce, per-tensor clip + Adam on head and trunk) AND value pass (the same with the value head), i.e.
every sample goes through both passes like `PPOAgent.update` (rl/agents/ppo.py:190-226); GAE / returns for
all bs x T transitions run once per update -- and once inside every timed region.  Rollout tensors (bs x T samples,
uint8 frames) are resident in HBM before the clock starts; every step gathers its minibatch from them.

`e2e` is the same metric through the reference-facing API with HOST buffers: pinned host transitions ->
`CARLAMemory.append` (the host->device copies) -> `CARLAgent.end_episode` (returns / GAE) -> `CARLAgent.update()`
(rl/agents/ppo.py:190-226: index batches, gather, both passes, clip + Adam) -> `write_summaries()` (the losses come
back to the host), all inside the clock.
"""
import argparse
import json
import re
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'carla-driving-rl-agent_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = 'ppo_update_samples_per_sec'
CONFIGS = {                                       # BASELINE.json `configs` (per-GPU minibatch 512 in all of them)
    'C2': dict(bs=512, T=256, height=90, width=120),      # batch_size=512, T=256, 90x120, bf16, 1 GPU  (C3 = the same on 8 GPUs)
    'C4': dict(bs=512, T=256, height=180, width=240),     # batch_size=1024 over 2 GPUs, 180x240 high-res tower
    'C5': dict(bs=512, T=512, height=90, width=120),      # batch_size=2048 over 4 GPUs, T=512 GAE / GRU stress
}
BYTES_PER_SAMPLE = {('bf16', 90, 120): 41.43e6, ('f32', 90, 120): 82.60e6, ('bf16', 180, 240): 164.33e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=12)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default=None, choices=sorted(CONFIGS), help='BASELINE.json configuration by name')
    ap.add_argument('--bs', type=int, default=512, help='samples per GPU per SGD step (BASELINE config 2: 512)')
    ap.add_argument('--T', type=int, default=256)
    ap.add_argument('--height', type=int, default=90)
    ap.add_argument('--width', type=int, default=120)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    ap.add_argument('--e2e-profile', action='store_true', help='print a torch.profiler table of one agent update (stderr)')
    ap.add_argument('--cpu-batch', type=int, default=32)
    a = ap.parse_args()
    if a.config:
        for k, v in CONFIGS[a.config].items():
            setattr(a, k, v)
    return a


# ----------------------------------------------------------------------------------------------------------------
def synthetic_rollout(torch, bs, T, H, W, device, seed):
    """Synthetic rollout tensors with the value distributions of SURVEY §8(d), generated on the device."""
    g = torch.Generator(device=device).manual_seed(seed)
    N = bs * T
    img = torch.empty(N, 4, H, W, 3, dtype=torch.uint8, device=device)
    chunk = 2048
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        img[s:e] = torch.randint(0, 256, (e - s, 4, H, W, 3), dtype=torch.uint8, device=device, generator=g)
    r = lambda *shape: torch.rand(*shape, device=device, generator=g)
    road = torch.cat([(r(N, 4, 3) < 0.2).float(), 0.3 + 0.6 * r(N, 4, 1),
                      torch.nn.functional.one_hot(torch.randint(0, 5, (N, 4), device=device, generator=g), 5).float()], -1)
    veh = torch.cat([r(N, 4, 1) * 2 - 1, r(N, 4, 3)], -1)
    nav = torch.sort(r(N, 4, 5) * 25, dim=-1).values
    d = dict(state_image=img, state_road=road.contiguous(), state_vehicle=veh.contiguous(), state_navigation=nav.contiguous(),
             actions=r(N, 2).clamp(1e-4, 1 - 1e-4), logp_old=0.5 * torch.randn(N, 2, device=device, generator=g),
             true_speed=0.3 * r(N, 1), true_sim=r(N, 1) * 2 - 1,
             rewards=(torch.randn(bs, T, device=device, generator=g) * 2 + 1).clamp(-10, 30),
             values_be=torch.stack([r(bs, T) * 2 - 1, r(bs, T) * 6], -1).contiguous(),
             last_be=torch.stack([r(bs) * 2 - 1, r(bs) * 6], -1).contiguous())
    return d


class ClockSampler(threading.Thread):
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_rate(steps, warmup, batch, H, W, threads=None):
    """Times the oracle's policy pass + value pass (+ clip + Adam) on `batch` samples per step on the host."""
    import torch
    from oracle import model, ppo, spec
    from tests import common as C
    if threads:
        torch.set_num_threads(threads)
    dyn, pol, val = C.fresh_params(torch.float32)
    obs, bt = C.synthetic_obs(batch, H, W, seed=1), C.synthetic_batch(batch, seed=2)
    st = {k: ({n: torch.zeros_like(t) for n, t in d.items()}, {n: torch.zeros_like(t) for n, t in d.items()})
          for k, d in (('dyn', dyn), ('pol', pol), ('val', val))}
    n_step = [0]

    def step():
        n_step[0] += 1
        r = C.policy_step_oracle(dyn, pol, obs, bt, dtype=torch.float32)
        ppo.apply_step(dyn, r['g_dyn'], st['dyn'][0], st['dyn'][1], 2 * n_step[0] - 1, 3e-4)
        ppo.apply_step(pol, r['g_head'], st['pol'][0], st['pol'][1], n_step[0], 3e-4, clip=1.0)
        r = C.value_step_oracle(dyn, val, obs, bt, dtype=torch.float32)
        ppo.apply_step(dyn, r['g_dyn'], st['dyn'][0], st['dyn'][1], 2 * n_step[0], 3e-4)
        ppo.apply_step(val, r['g_head'], st['val'][0], st['val'][1], n_step[0], 3e-4, clip=1.0)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads()


def workload_text(H, W, bs, world, T):
    return (f'PPO update (policy pass + value pass), obs {H}x{W}x3 x4-frame stack + road/vehicle/nav vectors, '
            f'bs={bs}/GPU (global {bs * world}), T={T}, N=bs*T rollout resident in HBM')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    # every host core, also under torchrun (which pins OMP_NUM_THREADS to 1 for its workers)
    rate, sec, cores = cpu_reference_rate(args.steps, max(1, min(args.warmup, 2)), args.cpu_batch, args.height, args.width,
                                          threads=os.cpu_count())
    sample = f'{args.cpu_batch}-sample SGD minibatch per step (policy pass + value pass + clip + Adam), fp32, torch-CPU restatement'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_text(args.height, args.width, args.bs, world, args.T),
                   'step': 'one SGD minibatch index through both passes; each timed step is a bounded sample of it (see cpu_baseline.sample)',
                   'note': 'TF 2.3.1 is not installable here; the CPU arm is the oracle port of the reference path (host cores, rank 0 only)'},
        'cpu_baseline': {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}), file=_REAL_STDOUT, flush=True)


# ---------------------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from cdra.engine import Engine
    from cdra.init import init_engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    debug = os.environ.get('BENCH_DEBUG')
    if debug:                                                       # where is a stuck rank?  (stderr, after BENCH_DEBUG seconds)
        import faulthandler
        faulthandler.dump_traceback_later(int(debug), exit=True)

    def stage(msg):
        if debug:
            print(f'[bench rank {rank}] {msg}', file=sys.stderr, flush=True)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
        stage('process group up')
    bs, T, H, W = args.bs, args.T, args.height, args.width
    eng = Engine(bs, H, W, dtype=args.dtype, image_u8=True, device=dev)
    init_engine(eng, seed=42)                                       # Keras-default random init, same on every rank
    old_pol = eng.pol.flat.clone()
    roll = synthetic_rollout(torch, bs, T, H, W, dev, 1234 + rank)
    N = bs * T
    perm_p = torch.randperm(N, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))
    perm_v = torch.randperm(N, device=dev, generator=torch.Generator(device=dev).manual_seed(8 + rank))
    keys = ('state_image', 'state_road', 'state_vehicle', 'state_navigation', 'actions', 'logp_old', 'true_speed', 'true_sim')
    mb = {k: torch.empty((bs,) + tuple(roll[k].shape[1:]), dtype=roll[k].dtype, device=dev) for k in keys}
    mb['adv'] = torch.empty(bs, device=dev); mb['returns'] = torch.empty(bs, 2, device=dev)
    gscale = 1.0 / world
    state = dict(adv=None, ret=None)

    def gae():
        ret, adv = eng.gae(roll['rewards'], roll['values_be'], roll['last_be'], 0.9999, 0.999, 2.0)
        state['ret'], state['adv'] = ret.view(N, 2), adv.view(N, 1)

    def gather(idx, what):              # the whole minibatch in one launch (cdra_gather_rows_multi), like CARLANetwork.gather_device
        extra = ('adv', 'adv') if what == 'policy' else ('ret', 'returns')
        eng.gather_rows_multi([roll[k] for k in keys] + [state[extra[0]]], idx, [mb[k] for k in keys] + [mb[extra[1]]])

    from cdra.parallel import GradSync
    sync = GradSync(eng)            # the library's own NCCL communicator: cdra_allreduce_grads on the compute stream

    def policy_pass(m):
        # rl/agents/ppo.py:199-210, core/carla_agent.py:351-388
        x = eng.dynamics_forward(m)
        eng.policy_head(x, m['actions'], m['logp_old'], m['adv'], m['true_speed'], m['true_sim'], 0.2, 1.0)
        eng.dynamics_backward(m, eng.d_x512)
        sync.allreduce_pass('policy')                                  # ONE collective: policy head + dynamics gradients
        eng.clip_adam('dyn', 3e-4, None, gscale)
        old_pol.copy_(eng.pol.flat)                                   # update_old_policy before the Adam step (ppo.py:249)
        eng.clip_adam('pol', 3e-4, 1.0, gscale)
        return eng.scalars[0].clone()

    def value_pass(m):
        # rl/agents/ppo.py:213-224, core/carla_agent.py:430-463
        x = eng.dynamics_forward(m)
        eng.value_head(x, m['returns'], m['true_speed'], m['true_sim'])
        eng.dynamics_backward(m, eng.d_x512)
        sync.allreduce_pass('value')
        eng.clip_adam('dyn', 3e-4, None, gscale)
        eng.clip_adam('val', 3e-4, 1.0, gscale)
        return eng.scalars[0]

    def sgd_step(i):
        """one SGD minibatch index through both passes; minibatches gathered from the HBM-resident rollout"""
        lo = (i % T) * bs
        gather(perm_p[lo:lo + bs], 'policy')
        loss_p = policy_pass(mb)
        gather(perm_v[lo:lo + bs], 'value')
        return loss_p, value_pass(mb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, first):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gae()                                                         # returns / GAE of all bs x T transitions: once per update, inside the clock
        for i in range(nsteps):
            sgd_step(first + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    launches0 = eng.lib.cdra_launch_count()
    stage('rollout ready')
    gae()
    for i in range(args.warmup):
        sgd_step(i)
        if debug:
            torch.cuda.synchronize(); stage(f'warm-up step {i} done')
    torch.cuda.synchronize()
    per_step_launches = (eng.lib.cdra_launch_count() - launches0) / max(1, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = eng.lib.cdra_launch_count()
    ms = timed(args.steps, args.warmup)
    launches = eng.lib.cdra_launch_count() - l0
    stage('timed region done')
    clocks = sampler.finish() if sampler else None
    # host time to ENQUEUE one step on an idle queue (no synchronisation inside): far below ms_per_step = the GPU, not the launch path, is the limit
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    sgd_step(args.warmup + args.steps)
    host_ms = 1e3 * (time.perf_counter() - h0)
    torch.cuda.synchronize()
    value = world * bs * args.steps / (ms / 1e3)

    peak, peak_src = measured_peaks()
    roof = None
    if not args.no_profile:
        # every rank runs the two profiled steps (they contain the gradient all-reduce); rank 0 reports
        roof = kernel_roofline(eng, sgd_step, args.warmup + args.steps, peak, peak_src)
    # ---- end-to-end: the same metric through the reference-facing API with HOST buffers (see e2e_agent_leg)
    e2e_steps = max(4, args.steps)                   # as many environment steps per update as timed SGD steps (amortises the per-update fixed work the same way)
    rollout_gb = roll['state_image'].numel() / 1e9
    del roll, mb
    torch.cuda.empty_cache()
    e2e = e2e_agent_leg(args, torch, dist, dev, world, rank, e2e_steps, barrier)

    out = {
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype,
        'data': 'synthetic',
        'config': {'workload': workload_text(H, W, bs, world, T),
                   'step': 'one SGD minibatch index = bs samples/GPU through both passes; returns/GAE of all bs x T transitions once inside the timed region',
                   'l2': f'inputs ({rollout_gb:.1f} GB rollout) and activations exceed L2; no flush needed',
                   'parallelism': f'dp{world}', 'optimizer': 'Keras-style Adam x3, per-tensor clip 1.0 on heads'},
        'e2e': e2e,
        'gpu_launches': int(launches), 'launches_per_step': per_step_launches,
        'host_enqueue_ms_per_step': host_ms, 'clocks': clocks,
    }
    bps = BYTES_PER_SAMPLE.get((args.dtype, H, W))
    if bps:
        out['roofline_step'] = {'bound': 'hbm', 'achieved': value / world * bps / 1e9, 'peak': peak, 'unit': 'GB/s',
                                'frac': value / world * bps / 1e9 / peak, 'bytes_per_sample': bps,
                                'note': 'canonical algorithmic bytes of SURVEY 8(d) x samples/s/GPU; peak ' + peak_src}
    if roof is not None and rank == 0:
        out['roofline'] = roof
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, sec, cores = cpu_reference_rate(12, 2, args.cpu_batch, H, W, threads=os.cpu_count())       # ~10 s of host work
        out['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                               'sample': f'12 timed SGD steps of {args.cpu_batch} samples (both passes + clip + Adam), fp32 oracle on the host'}
    if rank == 0:
        print(json.dumps(out), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_agent_leg(args, torch, dist, dev, world, rank, n_steps, barrier):
    """`e2e`: pinned HOST transitions -> CARLAMemory.append (H2D) -> end_episode (returns / GAE) -> CARLAgent.update()
    -> write_summaries() (losses D2H), timed as a whole with CUDA events; n_steps environment steps of bs parallel
    trajectories = n_steps SGD minibatch indices per pass (minibatch = bs samples)."""
    import tempfile
    from core import CARLAgent, SyntheticCARLAEnvironment
    bs, H, W = args.bs, args.height, args.width
    env = SyntheticCARLAEnvironment(image_shape=(H, W, 3), image_uint8=True, seed=rank)
    tmp = tempfile.mkdtemp(prefix='cdra_bench_')
    agent = CARLAgent(env, batch_size=bs, name='bench', weights_dir=os.path.join(tmp, 'w'), evaluation_dir=os.path.join(tmp, 'e'),
                      seed=42, skip_data=0, drop_batch_remainder=True, shuffle=True, shuffle_batches=False, log_mode='log',
                      policy_lr=3e-4, value_lr=3e-4, dynamics_lr=3e-4, entropy_regularization=1.0, clip_ratio=0.2, gamma=0.9999,
                      lambda_=0.999, advantage_scale=2.0, aug_intensity=0.0, network=dict(dtype=args.dtype, device=dev))
    g = torch.Generator().manual_seed(99 + rank)
    pin = lambda t: t.contiguous().pin_memory()
    r = lambda *shape: torch.rand(*shape, generator=g)
    steps = []
    for t in range(n_steps):                          # what bs parallel environments hand over at one step (host memory)
        state = dict(state_image=pin(torch.randint(0, 256, (bs, 4, H, W, 3), dtype=torch.uint8, generator=g)),
                     state_road=pin(torch.cat([(r(bs, 4, 3) < 0.2).float(), 0.3 + 0.6 * r(bs, 4, 1),
                                               torch.nn.functional.one_hot(torch.randint(0, 5, (bs, 4), generator=g), 5).float()], -1)),
                     state_vehicle=pin(torch.cat([r(bs, 4, 1) * 2 - 1, r(bs, 4, 3)], -1)),
                     state_navigation=pin(torch.sort(r(bs, 4, 5) * 25, dim=-1).values))
        steps.append(dict(state=state, action=pin(r(bs, 2).clamp(1e-4, 1 - 1e-4)), log_prob=pin(0.5 * torch.randn(bs, 2, generator=g)),
                          reward=pin((torch.randn(bs, generator=g) * 2 + 1).clamp(-10, 30)),
                          value=pin(torch.stack([r(bs) * 2 - 1, torch.floor(r(bs) * 6)], -1))))
    speed, sim = pin(r(n_steps * bs) * 30.0), pin(r(n_steps * bs) * 2 - 1)
    last = pin(torch.stack([r(bs) * 2 - 1, torch.floor(r(bs) * 6)], -1))
    h2d = sum(t.numel() * t.element_size() for st in steps for t in list(st['state'].values()) + [st['action'], st['log_prob'], st['reward'], st['value']])
    h2d += speed.numel() * 4 * 2 + last.numel() * 4
    mem = agent.get_memory(capacity=n_steps, num_envs=bs)            # preallocated device buffers (outside the clock)

    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))

    def run():
        del marks[:]
        mark('start')
        mem.delete()
        agent.memory = mem
        for st in steps:
            mem.append(st['state'], st['action'], st['reward'], st['value'], st['log_prob'])
        mark('append')
        env.info_buffer = dict(speed=speed.to(dev, non_blocking=True), similarity=sim.to(dev, non_blocking=True))
        agent.end_episode(last)
        mark('end_episode')
        agent.update()
        mark('update')
        agent.write_summaries()
        mark('summaries')
        return agent.statistics.last

    run()                                              # warm-up (allocations, first-use costs)
    if getattr(args, 'e2e_profile', False) and rank == 0:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            run()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=60), file=sys.stderr)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lastv = run()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    if rank == 0:                                      # where the end-to-end time goes (device time | host time per phase)
        torch.cuda.synchronize()
        for (n0, ev0, h0), (n1, ev1, h1) in zip(marks[:-1], marks[1:]):
            print(f'[e2e] {n1:12s} device {ev0.elapsed_time(ev1):8.2f} ms   host {1e3 * (h1 - h0):8.2f} ms', file=sys.stderr)
    assert all(k in lastv and lastv[k] == lastv[k] for k in ('loss_total', 'loss_value')), lastv
    return {'value': world * bs * n_steps / (ms / 1e3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d // n_steps,
            'd2h_bytes_per_step': max(4, agent.statistics.last_d2h_bytes // n_steps), 'ms_per_step': ms / n_steps, 'steps': n_steps,
            'note': 'CARLAMemory.append (pinned host -> device buffers) + end_episode (GAE) + CARLAgent.update() + write_summaries() '
                    'inside the clock; every sample crosses PCIe once and is used by both passes'}


def _demangle_short(name):
    """_ZN4cdra2v213dw_bwd_kernelILi120ELi1EEEv... -> dw_bwd<120,1> (length-prefixed Itanium identifiers)"""
    if not name.startswith('_ZN'):
        return name
    i, ident = 3, None
    while i < len(name) and name[i].isdigit():
        j = i
        while name[j].isdigit():
            j += 1
        n = int(name[i:j]); ident = name[j:j + n]; i = j + n
    if not ident:
        return name
    ints = []
    if i < len(name) and name[i] == 'I':
        ints = re.findall(r'L[ib](\d+)E', name[i:name.find('Ev', i) + 1 if name.find('Ev', i) > 0 else len(name)])
    ident = ident[:-7] if ident.endswith('_kernel') else ident
    return ident + ('<' + ','.join(ints) + '>' if ints else '')


def kernel_roofline(eng, sgd_step, first, peak, peak_src):
    """Per-kernel CUDA-event timing on the launch stream (serialised; outside the timed region)."""
    import ctypes
    import torch
    lib = eng.lib
    lib.cdra_profile_reset(); lib.cdra_profile_enable(1)
    sgd_step(first); sgd_step(first + 1)
    torch.cuda.synchronize()
    lib.cdra_profile_enable(0)
    n = lib.cdra_profile_report(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    lib.cdra_profile_report(buf, n + 16)
    rows = []
    for line in buf.value.decode().splitlines():
        name, cnt, ms, by = line.split('\t')
        short = name
        for tag in ('stem_fwd', 'pool_fwd', 'pw_fwd', 'dw_fwd', 'pass_fwd', 'gap_fwd', 'bstat', 'pw_dgrad', 'pw_wgrad', 'dw_dgrad',
                    'dw_wgrad', 'pass_bwd', 'pool_bwd', 'stem_wgrad', 'gap_bwd', 'sgemm', 'colsum', 'bn1d_fwd', 'bn1d_bwd',
                    'swish6_fwd', 'swish6_bwd', 'gru_gate_fwd', 'gru_gate_bwd', 'featnet_fwd', 'featnet_bwd', 'head_loss', 'gae',
                    'sqnorm', 'adam', 'gather_rows'):
            if tag in name:
                short = tag
        if short == name:                            # mangled name of a templated kernel: keep the kernel name + template integers
            short = _demangle_short(name)
        rows.append((short, int(cnt), float(ms), float(by)))
    agg = {}
    for s, c, m, b in rows:
        a = agg.setdefault(s, [0, 0.0, 0.0]); a[0] += c; a[1] += m; a[2] += b
    total = sum(a[1] for a in agg.values())
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])
    for k, v in top:                                 # full table (two steps) for the logs
        print(f'[profile] {k:32s} n={v[0]:4d} ms/step={v[1] / 2:7.3f} us/launch={1e3 * v[1] / max(v[0], 1):7.1f} '
              f'GB/s={(v[2] / (v[1] / 1e3) / 1e9) if v[1] > 0 and v[2] > 0 else 0:7.1f}', file=sys.stderr)
    name, (cnt, ms, by) = top[0]
    ach = by / (ms / 1e3) / 1e9 if ms > 0 else 0.0
    traffic = None                                   # DRAM bytes per launch of that kernel from the committed ncu capture
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
        if tj.get('kernel') == name:
            traffic = tj['dram_bytes_per_launch']
    except (OSError, ValueError, KeyError):
        pass
    return {'bound': 'hbm', 'kernel': name, 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': traffic,
            'algorithmic_bytes_per_launch': by / max(cnt, 1),
            'launches': cnt, 'avg_ms': ms / max(cnt, 1), 'share_of_step': ms / total if total else None, 'peak_source': peak_src,
            'by_kernel': {k: {'launches': v[0], 'ms': round(v[1], 3), 'share': round(v[1] / total, 4),
                              'GBps': round(v[2] / (v[1] / 1e3) / 1e9, 1) if v[1] > 0 and v[2] > 0 else None} for k, v in top[:20]}}


def _quiet_stdout():
    """stdout carries exactly ONE JSON line: library chatter written to fd 1 (e.g. NCCL's version banner) goes to stderr"""
    real = os.fdopen(os.dup(1), 'w')
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


if __name__ == '__main__':
    a = parse()
    _REAL_STDOUT = _quiet_stdout()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
