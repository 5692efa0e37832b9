"""N > 1 path on CPU: two gloo ranks, each with its own shard of the minibatch, exchange gradients exactly like the
NCCL path does on GPUs (cdra/parallel.py).  The kernels run through the logic-check build."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import PKG, ROOT

B, H, W = 2, 42, 58


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tests.emu.engine import EmuEngine
    from cdra.init import init_engine
    from cdra.parallel import GradSync
    from tests import common as C
    torch.set_num_threads(1)
    eng = EmuEngine(B, H, W, dtype='f32', image_u8=True, device='cpu')
    init_engine(eng, seed=100 + rank)                       # deliberately different: broadcast must fix it
    sync = GradSync(eng)
    sync.broadcast_parameters(0)
    obs, bt = C.synthetic_obs(B, H, W, seed=50 + rank), C.synthetic_batch(B, seed=60 + rank)   # rank-local shard
    C.policy_step_engine(eng, obs, bt)
    local = [eng.g_dyn.clone(), eng.g_pol.clone()]
    gathered = [[torch.zeros_like(t) for _ in range(world)] for t in local]
    for t, g in zip(local, gathered):
        dist.all_gather(g, t)
    sync.allreduce_pass('policy')              # ONE collective over the contiguous [policy | dynamics] gradient range
    assert torch.allclose(eng.g_dyn, sum(gathered[0]), atol=1e-6) and torch.allclose(eng.g_pol, sum(gathered[1]), atol=1e-6)
    eng.clip_adam('dyn', 3e-4, None, sync.grad_scale)
    eng.clip_adam('pol', 3e-4, 1.0, sync.grad_scale)
    torch.save(dict(dyn=eng.dyn.flat.clone(), pol=eng.pol.flat.clone(), norms=eng.norms['pol'].clone(),
                    mean_g=(sum(gathered[1]) / world)), os.path.join(out_dir, f'r{rank}.pt'))
    # the number of minibatches (= gradient all-reduces) of an update is agreed on: the shortest rank decides
    assert sync.agree_min(5 + 3 * rank) == 5
    lo, hi = sync.shard(8)
    assert (lo, hi) == (rank * 4, rank * 4 + 4)
    dist.destroy_process_group()


def test_two_rank_gradient_exchange(built_libs, tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f'r{i}.pt') for i in range(2))
    # replicas stay bit-identical after the step
    assert torch.equal(r0['dyn'], r1['dyn']) and torch.equal(r0['pol'], r1['pol'])
    # the per-tensor clip saw the globally averaged gradient: norms are those of mean_g, tensor by tensor
    from tests.emu.engine import EmuEngine
    eng = EmuEngine(B, H, W, dtype='f32', image_u8=True, device='cpu')
    want = torch.stack([eng.pol.view(n, r0['mean_g']).pow(2).sum() for n in eng.pol.names])
    assert torch.allclose(r0['norms'], want, rtol=1e-4, atol=1e-10)
