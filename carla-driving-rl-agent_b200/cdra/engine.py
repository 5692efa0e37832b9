"""Thin object layer over the C ABI: owns the plan, the flat parameter / state / gradient / Adam arenas
(torch tensors: device memory only) and the workspace, and exposes one method per ABI entry point.

`core.networks.CARLANetwork` / `core.carla_agent.CARLAgent` (the reference-facing API) are built on it.
"""
import ctypes as C
import torch

from . import _lib


class Arena:
    """Flat fp32 buffer + named views, mirroring Keras get_weights()/set_weights() bookkeeping."""

    def __init__(self, lib, plan, which, device):
        self.which = which
        self.size = int(lib.cdra_arena_size(plan, which))
        self.flat = torch.zeros(self.size, dtype=torch.float32, device=device)
        self.names, self.offsets, self.shapes = [], [], []
        n = lib.cdra_arena_num_tensors(plan, which)
        name = C.create_string_buffer(128)
        off, nd, dims = C.c_int64(), C.c_int32(), (C.c_int32 * 4)()
        for i in range(n):
            _lib.check(lib, lib.cdra_arena_tensor(plan, which, i, name, 128, C.byref(off), C.byref(nd), dims), 'arena_tensor')
            self.names.append(name.value.decode())
            self.offsets.append(off.value)
            self.shapes.append(tuple(dims[j] for j in range(nd.value)))
        self.index = {n_: i for i, n_ in enumerate(self.names)}
        self.offsets_dev = torch.tensor(self.offsets + [self.size], dtype=torch.int64, device=device)

    def view(self, name, flat=None):
        i = self.index[name]
        flat = self.flat if flat is None else flat
        n = 1
        for s in self.shapes[i]:
            n *= s
        return flat[self.offsets[i]:self.offsets[i] + n].view(self.shapes[i])

    def items(self, flat=None):
        return [(n, self.view(n, flat)) for n in self.names]

    def load_dict(self, d, strict=True):
        for n in self.names:
            if n in d:
                self.view(n).copy_(d[n].to(self.flat.dtype))
            elif strict:
                raise KeyError(n)

    def to_dict(self, flat=None):
        return {n: v.clone() for n, v in self.items(flat)}


class Engine:
    def __init__(self, batch, height=90, width=120, dtype='bf16', image_u8=True, device='cuda', share=None):
        """`share`: another Engine whose parameter / state / gradient / Adam arenas this one reuses (a plan is
        specific to one batch size; siblings let one network serve several, e.g. a remainder minibatch or B=1
        rollout inference)."""
        self.lib = self._open_library()
        self.device = torch.device(device)
        self._check_device()
        self.B, self.H, self.W = batch, height, width
        self.dtype = dtype
        self.image_u8 = bool(image_u8)
        cfg = _lib.Config(batch, height, width, _lib.BF16 if dtype == 'bf16' else _lib.F32, 1 if image_u8 else 0)
        self.plan = C.c_void_p()
        _lib.check(self.lib, self.lib.cdra_plan_create(C.byref(cfg), C.byref(self.plan)), 'plan_create')
        self.ws_bytes = int(self.lib.cdra_plan_workspace_bytes(self.plan))
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=self.device)
        if share is not None:
            for k in ('dyn', 'dyn_state', 'pol', 'pol_state', 'val', 'val_state', 'g_all', 'g_dyn', 'g_pol', 'g_val', 'g_policy_pass',
                      'g_value_pass', 'adam', 'adam_step', 'norms'):
                setattr(self, k, getattr(share, k))
        else:
            mk = lambda w: Arena(self.lib, self.plan, w, self.device)
            self.dyn, self.dyn_state = mk(_lib.ARENA_DYN_PARAMS), mk(_lib.ARENA_DYN_STATE)
            self.pol, self.pol_state = mk(_lib.ARENA_POL_PARAMS), mk(_lib.ARENA_POL_STATE)
            self.val, self.val_state = mk(_lib.ARENA_VAL_PARAMS), mk(_lib.ARENA_VAL_STATE)
            z = lambda a: torch.zeros_like(a.flat)
            # ONE flat gradient buffer [policy | dynamics | value] (each part padded to 64 floats): the policy pass's
            # gradients (policy + dynamics) and the value pass's (dynamics + value) are each one contiguous range, i.e.
            # one all-reduce per pass under data parallelism (cdra_allreduce_grads)
            r64 = lambda n: (n + 63) // 64 * 64
            o_dyn, o_val = r64(self.pol.size), r64(self.pol.size) + r64(self.dyn.size)
            self.g_all = torch.zeros(o_val + r64(self.val.size), dtype=torch.float32, device=self.device)
            self.g_pol = self.g_all[:self.pol.size]
            self.g_dyn = self.g_all[o_dyn:o_dyn + self.dyn.size]
            self.g_val = self.g_all[o_val:o_val + self.val.size]
            self.g_policy_pass = self.g_all[:o_dyn + self.dyn.size]          # what the policy pass exchanges
            self.g_value_pass = self.g_all[o_dyn:o_val + self.val.size]      # what the value pass exchanges
            self.adam = {k: (z(a), z(a)) for k, a in (('dyn', self.dyn), ('pol', self.pol), ('val', self.val))}
            self.adam_step = dict(dyn=0, pol=0, val=0)
            self.norms = {k: torch.zeros(len(a.names), dtype=torch.float32, device=self.device)
                          for k, a in (('pol', self.pol), ('val', self.val), ('dyn', self.dyn))}
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.x512, self.d_x512 = f(batch, 512), f(batch, 512)
        self.scalars, self.head_out = f(16), f(batch, 8)

    # the two places that tie an Engine to the GPU build of the library (the CPU test-suite subclasses them away)
    def _open_library(self):
        return _lib.load()

    def _check_device(self):
        if self.device.type != 'cuda':
            raise _lib.CdraError('libcdra runs on CUDA devices only (no CPU fallback)')

    def sibling(self, batch):
        return type(self)(batch, self.H, self.W, dtype=self.dtype, image_u8=self.image_u8, device=self.device, share=self)

    def __del__(self):
        try:
            if self.plan:
                self.lib.cdra_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def tensor(self, name):
        """View of a named intermediate tensor inside the workspace (parity taps)."""
        off, dims, es = C.c_int64(), (C.c_int32 * 4)(), C.c_int32()
        if self.dtype == 'bf16':
            # perf mode stores padded, shuffled channel planes: ask the library for the logical layout (fp32 copy)
            if self.lib.cdra_debug_export(self.plan, name.encode(), None, None, dims, None) == 0:
                out = torch.empty([d for d in dims], dtype=torch.float32, device=self.device)
                _lib.check(self.lib, self.lib.cdra_debug_export(self.plan, name.encode(), _lib.ptr(self.ws), _lib.ptr(out),
                                                                dims, self._stream()), 'debug_export')
                return out
        _lib.check(self.lib, self.lib.cdra_plan_tensor(self.plan, name.encode(), C.byref(off), dims, C.byref(es)), 'plan_tensor')
        shape = [d for d in dims]
        while len(shape) > 1 and shape[-1] == 1:
            shape.pop()
        n = 1
        for s in shape:
            n *= s
        dt = torch.float32 if es.value == 4 else torch.bfloat16
        raw = self.ws[off.value: off.value + n * es.value]
        return raw.view(dt).view(shape)

    def _check_obs(self, obs):
        img = obs['state_image']
        want = torch.uint8 if self.image_u8 else torch.float32
        assert img.dtype == want and tuple(img.shape) == (self.B, 4, self.H, self.W, 3), (img.dtype, img.shape)
        for k, d in (('state_road', 9), ('state_vehicle', 4), ('state_navigation', 5)):
            assert obs[k].dtype == torch.float32 and tuple(obs[k].shape) == (self.B, 4, d), (k, obs[k].shape)
            assert obs[k].is_contiguous()
        assert img.is_contiguous()

    # ------------------------------------------------------------------ ABI calls
    def dynamics_forward(self, obs, training=True, out=None):
        self._check_obs(obs)
        out = self.x512 if out is None else out
        p = _lib.ptr
        _lib.check(self.lib, self.lib.cdra_dynamics_forward(
            self.plan, p(self.dyn.flat), p(self.dyn_state.flat), p(obs['state_image']), p(obs['state_road']),
            p(obs['state_vehicle']), p(obs['state_navigation']), 1 if training else 0, p(out), p(self.ws),
            self._stream()), 'dynamics_forward')
        return out

    def dynamics_backward(self, obs, d_out):
        p = _lib.ptr
        _lib.check(self.lib, self.lib.cdra_dynamics_backward(
            self.plan, p(self.dyn.flat), p(obs['state_image']), p(obs['state_road']), p(obs['state_vehicle']),
            p(obs['state_navigation']), p(d_out), p(self.g_dyn), p(self.ws), self._stream()), 'dynamics_backward')
        return self.g_dyn

    def debug_stem_backward(self, obs, legacy):
        """stem gradients only, from the workspace a training step left behind (parity aid) -> fresh gradient arena"""
        g = torch.zeros_like(self.g_dyn)
        _lib.check(self.lib, self.lib.cdra_debug_stem_backward(
            self.plan, _lib.ptr(self.dyn.flat), _lib.ptr(obs['state_image']), _lib.ptr(g), _lib.ptr(self.ws),
            1 if legacy else 0, self._stream()), 'debug_stem_backward')
        return g

    def policy_head(self, x512, actions_eval, logp_old, adv, true_speed, true_sim, clip_ratio=0.2, ent_coef=1.0,
                    training=True, grad_scale=1.0, backward=True, update_moving=True, actions_jac=None):
        p = _lib.ptr
        state = self.pol_state.flat if (update_moving or not training) else None
        _lib.check(self.lib, self.lib.cdra_policy_head_loss_fwd_bwd(
            self.plan, p(self.pol.flat), p(state), p(x512), p(actions_eval), p(actions_jac), p(logp_old), p(adv),
            p(true_speed), p(true_sim), clip_ratio, ent_coef, 1 if training else 0, grad_scale, p(self.scalars),
            p(self.head_out), p(self.d_x512) if backward else None, p(self.g_pol) if backward else None, p(self.ws),
            self._stream()), 'policy_head')
        return self.scalars

    def value_head(self, x512, returns_be, true_speed, true_sim, training=True, grad_scale=1.0, backward=True):
        p = _lib.ptr
        _lib.check(self.lib, self.lib.cdra_value_head_loss_fwd_bwd(
            self.plan, p(self.val.flat), p(self.val_state.flat), p(x512), p(returns_be), p(true_speed), p(true_sim),
            1 if training else 0, grad_scale, p(self.scalars), p(self.head_out), p(self.d_x512) if backward else None,
            p(self.g_val) if backward else None, p(self.ws), self._stream()), 'value_head')
        return self.scalars

    def gae(self, rewards, values_be, last_value_be, gamma, lambda_, scale):
        bs, T = rewards.shape
        returns_be = torch.empty(bs, T, 2, dtype=torch.float32, device=rewards.device)
        adv = torch.empty(bs, T, dtype=torch.float32, device=rewards.device)
        p = _lib.ptr
        _lib.check(self.lib, self.lib.cdra_gae(p(rewards), p(values_be), p(last_value_be), float(gamma), float(lambda_),
                                               float(scale), bs, T, p(returns_be), p(adv), self._stream()), 'gae')
        return returns_be, adv

    def clip_adam(self, which, lr, clip_norm=None, grad_scale=1.0, beta1=0.9, beta2=0.999, eps=1e-7):
        arena = dict(dyn=self.dyn, pol=self.pol, val=self.val)[which]
        grads = dict(dyn=self.g_dyn, pol=self.g_pol, val=self.g_val)[which]
        m, v = self.adam[which]
        self.adam_step[which] += 1
        p = _lib.ptr
        _lib.check(self.lib, self.lib.cdra_clip_adam(
            p(arena.flat), p(grads), p(m), p(v), p(arena.offsets_dev), len(arena.names), arena.size,
            float(clip_norm) if clip_norm else 0.0, float(lr), beta1, beta2, eps, self.adam_step[which],
            float(grad_scale), p(self.norms[which]), self._stream()), 'clip_adam')

    def grad_norms(self, which, grad_scale=1.0):
        """per-tensor L2 norms of a gradient arena as ONE device tensor (one launch, no host synchronisation)"""
        arena = dict(dyn=self.dyn, pol=self.pol, val=self.val)[which]
        grads = dict(dyn=self.g_dyn, pol=self.g_pol, val=self.g_val)[which]
        out = torch.empty(len(arena.names), dtype=torch.float32, device=self.device)
        p = _lib.ptr
        _lib.check(self.lib, self.lib.cdra_grad_norms(p(grads), p(arena.offsets_dev), len(arena.names), arena.size,
                                                      float(grad_scale), p(out), self._stream()), 'grad_norms')
        return out.sqrt_()

    def gather_rows_multi(self, srcs, index, outs):
        """dst[k][i] = src[k][index[i]] for every tensor of a minibatch in one launch (same index vector)"""
        n = len(srcs)
        if n > 12:
            for a, b in ((srcs[:12], outs[:12]), (srcs[12:], outs[12:])):
                self.gather_rows_multi(a, index, b)
            return outs
        PA = C.c_void_p * n
        sp = PA(*[t.data_ptr() for t in srcs]); dp = PA(*[t.data_ptr() for t in outs])
        rb = (C.c_int64 * n)(*[t[0].numel() * t.element_size() for t in srcs])
        _lib.check(self.lib, self.lib.cdra_gather_rows_multi(sp, dp, rb, n, _lib.ptr(index), index.numel(), self._stream()),
                   'gather_rows_multi')
        return outs

    def gather_rows(self, src, index, out):
        row_bytes = src[0].numel() * src.element_size()
        p = _lib.ptr
        _lib.check(self.lib, self.lib.cdra_gather_rows(p(src), p(index), index.numel(), row_bytes, p(out),
                                                       self._stream()), 'gather_rows')
        return out
