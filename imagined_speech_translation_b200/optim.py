"""Fused optimizer step for the train loop (flat fp32 buffers, two HBM-bound kernels per group).

Mirrors what ``EEGTrainer.train_epoch`` does every ``accumulation_steps`` micro-batches
(reference ``main_model/src/training/trainer.py:101-113``): ``clip_grad_norm_`` over *all*
parameters, then ``AdamW.step()`` with the three learning-rate groups of
``config/training_config.py:55-77``.  Update rule: ``torch.optim.AdamW`` (SURVEY.md section 7:
the reference's ``transformers.AdamW`` no longer exists; parity at the optimizer is pinned
against torch's).  Parameters that never receive a gradient (the BART *encoder*, 43.3 M
parameters) are left untouched, exactly as torch skips ``grad is None``.

Subclasses ``torch.optim.Optimizer`` so the reference's LR schedulers
(``get_cosine_schedule_with_warmup``) drive ``param_groups[i]['lr']`` unchanged.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, nn_ops


class ParamStack:
    """G same-shaped parameters (one per region encoder) that FlatAdamW has laid out back to back: ``p`` /
    ``g`` are (G, *shape) fp32 views of the flat master / gradient buffers, ``w16`` the (G, *shape) bf16 shadow
    the AdamW kernel keeps current.  The lock-step region path (``grouped.py``) reads its weights and
    accumulates its weight gradients through these views -- one launch serves all G modules."""

    def __init__(self, params, p, g, w16):
        self.params, self.p, self.g, self._w16 = list(params), p, g, w16

    def __len__(self):
        return len(self.params)

    def bound(self) -> bool:
        """True while every parameter still is its slice of the stacked buffers."""
        step = self.p.stride(0) * 4
        base, gbase = self.p.data_ptr(), self.g.data_ptr()
        return all(q.data_ptr() == base + i * step and q.grad is not None and q.grad.data_ptr() == gbase + i * step
                   for i, q in enumerate(self.params))

    def w16(self) -> torch.Tensor:
        """The bf16 shadow, refreshed first if a parameter was modified in place since the last AdamW step."""
        if any(getattr(q, "_eegx_w16_ver", -1) != q._version for q in self.params):
            with torch.no_grad():
                self._w16.copy_(self.p)
            for q in self.params:
                q._eegx_w16_ver = q._version
        return self._w16


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, stacks=None):
        """stacks: optional list of parameter lists (e.g. ``BrainRegionEncoder.parameter_stacks()``): the members of
        a list have one shape and are laid out back to back in the flat buffers, so that ``ParamStack`` views exist."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._stack_request = [list(st) for st in (stacks or []) if len(st) > 1]
        self._flat = None          # per group: dict(p, g, m, v, params)
        self._step = 0
        self._norm_sq = None
        self._ws = None
        self.grad_scale = 1.0      # e.g. 1 / world_size when the flat grads hold an all-reduced SUM

    # ------------------------------------------------------------------ flat buffers
    def _build(self):
        """One contiguous fp32 gradient buffer for ALL groups (a single all-reduce covers it) and one
        parameter / moment buffer per group (the groups differ in learning rate)."""
        plan = []
        live_stacks = []
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.grad is not None]
            for p in ps:
                if p.dtype != torch.float32 or not p.is_cuda:
                    raise _lib.EegxError("FlatAdamW needs float32 CUDA parameters (no CPU fallback)")
            # members of a requested stack follow its first member (same shape, all in this group, all with gradients,
            # 8-element multiples so that the stacked views are contiguous)
            in_group = {id(p) for p in ps}
            follow, skip = {}, set()
            for st in self._stack_request:
                if (all(id(q) in in_group for q in st) and len({tuple(q.shape) for q in st}) == 1
                        and st[0].numel() % 8 == 0 and not any(id(q) in skip or id(q) in follow for q in st)):
                    follow[id(st[0])] = st
                    skip.update(id(q) for q in st[1:])
                    live_stacks.append(st)
            ordered = []
            for p in ps:
                if id(p) in skip:
                    continue
                ordered.extend(follow.get(id(p), [p]))
            ps = ordered
            sizes = [(p.numel() + 7) // 8 * 8 for p in ps]          # every fp32 AND bf16-shadow view 16-byte aligned
            plan.append((ps, sizes))
        if not any(ps for ps, _ in plan):
            raise _lib.EegxError("FlatAdamW.step() called before any backward pass")
        dev = next(ps[0].device for ps, _ in plan if ps)
        n_grads = sum(sum(sz) for _, sz in plan)
        from . import distributed as _dp
        # in a multi-rank NCCL job the gradient buffer comes from NCCL's allocator and is registered with the
        # communicator (zero-copy / NVLS all-reduce on the buffer itself); otherwise plain device memory
        self._all_grads = _dp.nccl_registered_zeros(n_grads, dev)
        if self._all_grads is None:
            self._all_grads = torch.zeros(n_grads, device=dev)
        flats, base = [], 0
        self._offsets = {}         # id(param) -> (offset, padded length) inside _all_grads
        for ps, sizes in plan:
            if not ps:
                flats.append(None)
                continue
            total = sum(sizes)
            flat_p = torch.zeros(total, device=dev)
            flat_w16 = torch.zeros(total, device=dev, dtype=torch.bfloat16)     # bf16 shadow, kept current by the AdamW kernel
            flat_g = self._all_grads[base:base + total]
            off = 0
            for p, n in zip(ps, sizes):
                k = p.numel()
                self._offsets[id(p)] = (base + off, n)
                flat_p[off:off + k].copy_(p.data.reshape(-1))
                flat_g[off:off + k].copy_(p.grad.reshape(-1))
                p.data = flat_p[off:off + k].view_as(p)
                p.grad = flat_g[off:off + k].view_as(p)
                p._eegx_w16 = flat_w16[off:off + k].view_as(p)
                p._eegx_w16_ver = p._version                  # flat_w16 is filled from flat_p right below
                p._eegx_flat = (len(flats), off)               # (group index, offset inside the group's buffers)
                off += n
            base += total
            flat_w16.copy_(flat_p)
            flats.append(dict(p=flat_p, g=flat_g, m=torch.zeros_like(flat_p), v=torch.zeros_like(flat_p),
                              w16=flat_w16, params=ps))
        self._flat = flats
        self.stacks = []
        for st in live_stacks:
            gi, off = st[0]._eegx_flat
            f, k, G = flats[gi], st[0].numel(), len(st)
            shape = (G,) + tuple(st[0].shape)
            stack = ParamStack(st, f['p'][off:off + G * k].view(shape), f['g'][off:off + G * k].view(shape),
                               f['w16'][off:off + G * k].view(shape))
            for i, q in enumerate(st):
                q._eegx_stack = (stack, i)
            self.stacks.append(stack)
        # parameters left outside (no gradient at build time) and the addresses the views must keep: step() checks both
        self._outside = [p for g in self.param_groups for p in g['params'] if id(p) not in self._offsets]
        self._expected_ptr = [(p, p.data_ptr(), p.grad.data_ptr()) for f in flats if f is not None for p in f['params']]
        pending, self._pending_state = getattr(self, "_pending_state", None), None
        if pending is not None:
            self._import_state(pending)
        self._norm_sq = torch.zeros(1, device=dev)
        self._ws = torch.empty(_lib.lib().eegx_sumsq_workspace_bytes(), dtype=torch.uint8, device=dev)

    def flat_grads(self):
        """The flat gradient buffers (one per group that has gradients): the data-parallel
        all-reduce runs directly on these."""
        if self._flat is None:
            self._build()
        return [self._all_grads]

    def grad_runs(self, params):
        """Contiguous [lo, hi) runs of the flat gradient buffer that hold the gradients of ``params``
        (parameters without a gradient are skipped), merged where adjacent."""
        if self._flat is None:
            self._build()
        spans = sorted(self._offsets[id(p)] for p in params if id(p) in self._offsets)
        runs = []
        for off, n in spans:
            if runs and runs[-1][1] == off:
                runs[-1][1] = off + n
            else:
                runs.append([off, off + n])
        return [tuple(r) for r in runs]

    def zero_grad(self, set_to_none: bool = False):
        if self._flat is None:
            return super().zero_grad(set_to_none=True)
        self._all_grads.zero_()

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        if closure is not None:
            raise ValueError("closure is not supported")
        if self._flat is None:
            self._build()
        self._check_bindings()
        lib = _lib.lib()
        st = _lib.stream_ptr()
        self._step += 1
        norm_ptr = None
        if max_grad_norm is not None:
            first = True
            for f in self._flat:
                if f is None:
                    continue
                _lib.check(lib.eegx_sumsq_f32(_lib.ptr(f['g']), f['g'].numel(), _lib.ptr(self._norm_sq),
                                              0 if first else 1, _lib.ptr(self._ws), self._ws.numel(), st),
                           "eegx_sumsq_f32")
                first = False
            norm_ptr = _lib.ptr(self._norm_sq)
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            b1, b2 = group['betas']
            _lib.check(lib.eegx_adamw_clip_f32(
                _lib.ptr(f['p']), _lib.ptr(f['g']), _lib.ptr(f['m']), _lib.ptr(f['v']), f['p'].numel(),
                float(group['lr']), float(b1), float(b2), float(group['eps']), float(group['weight_decay']),
                self._step, norm_ptr, float(max_grad_norm or 0.0), float(self.grad_scale), _lib.ptr(f['w16']), st),
                "eegx_adamw_clip_f32")
            for p in f['params']:
                p._eegx_w16_ver = p._version       # the shadow matches the parameter as of now
        nn_ops.clear_pack_cache()      # parameters changed behind autograd's back: derived packs are stale
        return None

    def _check_bindings(self):
        """The flat buffers are only correct while every parameter still IS its view of them: fail loudly when a
        parameter got its first gradient after the buffers were laid out (it would never be updated nor clipped),
        or when ``.data`` / ``.grad`` was rebound behind the optimizer's back (``model.to()``, ``load_state_dict``
        with ``assign=True``, ``zero_grad(set_to_none=True)`` by foreign code)."""
        for p in self._outside:
            if p.grad is not None:
                raise _lib.EegxError(
                    "FlatAdamW: a parameter received its first gradient after the flat buffers were built "
                    f"(shape {tuple(p.shape)}); run the first optimizer step on a batch that exercises every "
                    "trainable parameter, or rebuild the optimizer")
        for p, ptr, gptr in self._expected_ptr:
            if p.data_ptr() != ptr or p.grad is None or p.grad.data_ptr() != gptr:
                raise _lib.EegxError(
                    f"FlatAdamW: parameter of shape {tuple(p.shape)} no longer points into the flat buffers "
                    "(its .data or .grad was rebound after the first step); rebuild the optimizer")

    # ------------------------------------------------------------------ checkpoint interchange
    def state_dict(self):
        """``torch.optim.AdamW``'s layout: ``state[i] = {'step', 'exp_avg', 'exp_avg_sq'}`` per parameter index in
        ``param_groups`` order (parameters that never received a gradient have no entry, as in torch), so a
        checkpoint written here resumes under ``torch.optim.AdamW`` and vice versa (reference
        ``trainer.py:339-385`` stores ``optimizer.state_dict()`` under 'optimizer_state_dict')."""
        base = super().state_dict()
        if self._flat is None:
            if getattr(self, "_pending_state", None) is not None:
                base['state'] = self._pending_state
            return base
        index, i = {}, 0
        for g in self.param_groups:
            for p in g['params']:
                index[id(p)] = i
                i += 1
        state = {}
        for f in self._flat:
            if f is None:
                continue
            lo = self._offsets[id(f['params'][0])][0]
            for p in f['params']:
                off = self._offsets[id(p)][0] - lo
                k = p.numel()
                state[index[id(p)]] = {
                    'step': torch.tensor(float(self._step)),
                    'exp_avg': f['m'][off:off + k].view_as(p).clone(),
                    'exp_avg_sq': f['v'][off:off + k].view_as(p).clone(),
                }
        base['state'] = state
        return base

    def load_state_dict(self, state_dict):
        groups = state_dict['param_groups']
        if len(groups) != len(self.param_groups) or any(len(a['params']) != len(b['params'])
                                                        for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has different parameter groups")
        for mine, theirs in zip(self.param_groups, groups):
            mine.update({k: v for k, v in theirs.items() if k != 'params'})
        state = {int(k): v for k, v in state_dict.get('state', {}).items()}
        if self._flat is None:
            self._pending_state = state            # applied when the flat buffers are laid out (first step)
            steps = {int(float(v['step'])) for v in state.values()}
            self._step = max(steps) if steps else 0
            return
        self._import_state(state)

    def _import_state(self, state):
        params = [p for g in self.param_groups for p in g['params']]
        steps = set()
        for f in self._flat:
            if f is not None:
                f['m'].zero_()
                f['v'].zero_()
        pos = {id(p): i for i, p in enumerate(params)}
        for f in self._flat:
            if f is None:
                continue
            lo = self._offsets[id(f['params'][0])][0]
            for p in f['params']:
                st = state.get(pos[id(p)])
                if st is None:
                    continue
                off = self._offsets[id(p)][0] - lo
                k = p.numel()
                f['m'][off:off + k].copy_(st['exp_avg'].reshape(-1))
                f['v'][off:off + k].copy_(st['exp_avg_sq'].reshape(-1))
                steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise _lib.EegxError(f"FlatAdamW keeps one step counter; the loaded state has {sorted(steps)}")
        self._step = steps.pop() if steps else 0

    def grad_norm(self) -> torch.Tensor:
        """||g||_2 of the last clipped step (device scalar, no sync)."""
        return self._norm_sq.sqrt() * self.grad_scale
