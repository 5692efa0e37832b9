"""Data parallelism over the minibatch: one process per GPU, parameters / Adam state replicated,
each rank computes gradients on its shard of the global minibatch (BatchNorm statistics per replica, decision
D3), one all-reduce (sum) per gradient arena per SGD step over NCCL/NVLink, and the fused clip+Adam kernel applies
`grad_scale = 1 / world_size` while it reads the reduced arena — so per-tensor clipping sees the global-batch
gradient, exactly like tf.clip_by_norm on a single process (rl/utils.py:120-121).

torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests); there is no data-path collective besides
this exchange.
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class GradSync:
    """Gradient exchange of one engine.  On GPUs the all-reduce is the library's own `cdra_allreduce_grads` (an NCCL
    communicator created from an id that rank 0 broadcasts through torch.distributed), enqueued on the engine's stream like
    every kernel; on the CPU test path (gloo) it is `torch.distributed.all_reduce`."""

    def __init__(self, engine):
        self.engine = engine
        self.world = world_size()
        self.grad_scale = 1.0 / self.world
        self.comm = None
        if self.world > 1 and engine.device.type == 'cuda':
            import ctypes as C
            from . import _lib
            lib = engine.lib
            ident = torch.zeros(128, dtype=torch.uint8)
            if rank() == 0:
                buf = (C.c_ubyte * 128)()
                _lib.check(lib, lib.cdra_comm_unique_id(buf), 'comm_unique_id')
                ident = torch.tensor(list(buf), dtype=torch.uint8)
            ident = ident.to(engine.device)
            dist.broadcast(ident, 0)
            raw = (C.c_ubyte * 128)(*ident.cpu().tolist())
            comm = C.c_void_p()
            torch.cuda.set_device(engine.device)
            _lib.check(lib, lib.cdra_comm_create(raw, self.world, rank(), C.byref(comm)), 'comm_create')
            self.comm = comm

    def __del__(self):
        try:
            if self.comm:
                self.engine.lib.cdra_comm_destroy(self.comm)
                self.comm = None
        except Exception:
            pass

    def broadcast_parameters(self, src=0):
        """Make every replica start from rank `src`'s parameters, state and Adam moments."""
        if self.world == 1:
            return
        e = self.engine
        for t in (e.dyn.flat, e.dyn_state.flat, e.pol.flat, e.pol_state.flat, e.val.flat, e.val_state.flat):
            dist.broadcast(t, src)
        for m, v in e.adam.values():
            dist.broadcast(m, src); dist.broadcast(v, src)

    def _sum(self, t):
        if self.comm is not None:
            from . import _lib
            e = self.engine
            _lib.check(e.lib, e.lib.cdra_allreduce_grads(self.comm, _lib.ptr(t), t.numel(), e._stream()), 'allreduce_grads')
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def allreduce(self, *which):
        """Sum the named gradient arenas ('dyn', 'pol', 'val') across ranks, in place, on the current stream."""
        if self.world == 1:
            return
        e = self.engine
        for w in which:
            self._sum(dict(dyn=e.g_dyn, pol=e.g_pol, val=e.g_val)[w])

    def allreduce_pass(self, which):
        """ONE collective for everything a pass produced: 'policy' = policy head + dynamics, 'value' = dynamics + value head
        (contiguous ranges of the engine's flat gradient buffer)."""
        if self.world == 1:
            return
        e = self.engine
        self._sum(e.g_policy_pass if which == 'policy' else e.g_value_pass)

    def agree_min(self, value: int) -> int:
        """The smallest `value` over all ranks.  Every rank must issue the same number of gradient all-reduces: the number
        of minibatches of an update (and whether the update happens at all, core/carla_agent.py:130-133) is a rank-local
        quantity -- an early `done` shortens one rank's memory -- so it is agreed on before the loop."""
        if self.world == 1:
            return int(value)
        t = torch.tensor([int(value)], dtype=torch.int64, device=self.engine.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def shard(self, n_trajectories):
        """Trajectories [lo, hi) owned by this rank (rollouts never move between GPUs)."""
        per = n_trajectories // self.world
        r = rank()
        return r * per, (r + 1) * per
