"""Compute-dtype shadows of the fp32 master parameters.

The reference keeps fp32 ``nn.Parameter``s (state_dict contract).  In bf16 mode
the tensor-core GEMMs read bf16 copies.  Two mechanisms keep them fresh:

* per-parameter cache keyed on ``(id, _version, data_ptr)`` -- works for any module used stand-alone;
* :class:`FlatParams` -- re-homes every parameter of a model as a view of ONE flat fp32 buffer
  (and its ``.grad`` as a view of one flat gradient buffer), so a single ``dl_cast`` launch
  refreshes all shadows and a single NCCL all-reduce covers all gradients.
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional

import torch
from torch.utils.weak import WeakIdKeyDictionary

from . import kernels as K

# All registries hold the parameters weakly: entries die with the model, and a new parameter that
# happens to reuse a Python id or a device address can never hit a stale entry.
# (WeakIdKeyDictionary: identity-keyed -- tensors overload ==, so the stdlib weak dict cannot hold them)
_CACHE = WeakIdKeyDictionary()      # Parameter -> [version, ptr, shadow]
_FLAT = WeakIdKeyDictionary()       # Parameter -> weakref to its FlatParams store
_BY_PTR: Dict[int, "weakref.ref"] = {}                              # data_ptr of a flat view -> Parameter
_GROUPS: Dict[tuple, tuple] = {}     # data_ptrs of a fused parameter group -> (weakref store, offset, shape)


def flat_grad_of(t: torch.Tensor) -> Optional[torch.Tensor]:
    """The flat-buffer gradient view of the parameter `t` refers to (None when `t` is not a whole
    parameter of a FlatParams store).  Used to accumulate weight gradients in place."""
    r = _BY_PTR.get(t.data_ptr()) if t.dtype == torch.float32 else None
    p = r() if r is not None else None
    if p is None:
        if r is not None:
            _BY_PTR.pop(t.data_ptr(), None)
        return None
    if p.data_ptr() != t.data_ptr() or p.shape != t.shape or p.grad is None or p.grad.dtype != torch.float32:
        return None
    sr = _FLAT.get(p)
    store = sr() if sr is not None else None
    if store is not None and not store.inplace_grads:
        return None
    mark_touched(p)            # a kernel is about to accumulate this parameter's gradient in place
    return p.grad


def mark_touched(p: torch.Tensor) -> None:
    """Record that `p` receives a gradient in the current step (FlatAdamW updates only those)."""
    r = _BY_PTR.get(p.data_ptr())           # by device address: robust to re-wrapped tensors
    q = r() if r is not None else None
    if q is None:
        return
    r = _FLAT.get(q)
    fp = r() if r is not None else None
    if fp is not None:
        fp.touched.add(fp.index_of[id(q)])


def shadow(p: torch.Tensor) -> torch.Tensor:
    """`p` (fp32 parameter or a plain tensor) as a contiguous tensor in the compute dtype."""
    cd = K.compute_dtype()
    if p.dtype == cd:
        return p.detach()
    if not isinstance(p, torch.nn.Parameter):      # temporaries (slices of parameters): no caching
        return K.cast(p.detach().contiguous(), cd)
    r = _FLAT.get(p)
    fp = r() if r is not None else None
    if fp is not None:
        return fp.shadow_of(p)
    e = _CACHE.get(p)
    if e is not None and e[0] == p._version and e[1] == p.data_ptr() and e[2].dtype == cd:
        return e[2]
    t = e[2] if (e is not None and e[2].dtype == cd and e[2].shape == p.shape) else \
        torch.empty(p.shape, dtype=cd, device=p.device)
    K.cast(p.detach(), cd, out=t)
    _CACHE[p] = [p._version, p.data_ptr(), t]
    return t


class FlatParams:
    """All parameters of `module` as views of one flat fp32 buffer (+ flat grads, + bf16 shadow)."""

    ALIGN = 64  # elements; keeps every view 256-byte aligned (TMA needs 16 B)

    def __init__(self, module: torch.nn.Module, groups=None, first=None):
        """groups: lists of parameters to lay out back to back (e.g. the query / key / value weights
        of one attention block), so that :func:`fused_group` can hand a GEMM the concatenated
        [sum out_i, in] matrix, its gradient and its bf16 shadow as plain views -- three projections
        of one input become one launch forward and one per gradient."""
        params: List[torch.nn.Parameter] = []
        seen = set()
        for grp in (groups or []):
            for p in grp:
                if id(p) not in seen and p.dtype == torch.float32:
                    seen.add(id(p))
                    params.append(p)
        # `first`: parameters to place directly behind the groups, so that [0, head_numel) is one
        # contiguous range (the PMMA parameters: their gradients are complete long before the rest of
        # the backward, and that range is all-reduced while the rest still runs)
        for p in (first or []):
            if id(p) not in seen and p.dtype == torch.float32:
                seen.add(id(p))
                params.append(p)
        n_head = len(params)
        for p in module.parameters():
            if id(p) not in seen and p.dtype == torch.float32:
                seen.add(id(p))
                params.append(p)
        if not params:
            raise ValueError("module has no fp32 parameters")
        dev = params[0].device
        self.params = params
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        self.head_numel = self.offsets[n_head] if n_head < len(params) else off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.shadow16: Optional[torch.Tensor] = None
        self._synced = -1
        self._views16: Dict[int, torch.Tensor] = {}
        # Which parameters received a gradient since the last zero_grad(): torch's optimizers skip
        # parameters whose .grad is None, but here .grad always exists (a view of the flat buffer), so
        # the producers say so -- kernels that accumulate in place go through flat_grad_of(), and
        # gradients that arrive through autograd's AccumulateGrad fire the hook below.
        self.touched = set()
        # True: the backward kernels ADD weight / bias gradients straight into the flat gradient buffer
        # and hand autograd None for those inputs -- AccumulateGrad (and with it torch DDP's reducer
        # hooks, torch.autograd.grad, gradient hooks) never sees them.  TrainStep / ContrastiveStep
        # all-reduce the flat buffer themselves; under torch / Lightning DDP set this to False so every
        # parameter gradient travels through autograd as usual.
        self.inplace_grads = True
        self.index_of = {id(p): i for i, p in enumerate(params)}
        self._hooks = []
        with torch.no_grad():
            for i, (p, o) in enumerate(zip(params, self.offsets)):
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
                _FLAT[p] = weakref.ref(self)
                _BY_PTR[p.data_ptr()] = weakref.ref(p)
                if p.requires_grad:
                    self._hooks.append(p.register_post_accumulate_grad_hook(
                        lambda _p, i=i, t=self.touched: t.add(i)))
        # fusable groups: members adjacent without padding gaps, same trailing shape
        off_of = {id(p): o for p, o in zip(params, self.offsets)}
        self.groups: Dict[tuple, tuple] = {}
        for grp in (groups or []):
            ok = all(id(p) in off_of for p in grp) and len({tuple(p.shape[1:]) for p in grp}) == 1
            for a, b in zip(grp[:-1], grp[1:]):
                ok = ok and off_of[id(b)] == off_of[id(a)] + a.numel()
            if ok:
                rows = sum(p.shape[0] for p in grp)
                ent = (off_of[id(grp[0])], (rows,) + tuple(grp[0].shape[1:]))
                self.groups[tuple(id(p) for p in grp)] = ent
                _GROUPS[tuple(p.data_ptr() for p in grp)] = (weakref.ref(self),) + ent

    def zero_grad(self) -> None:
        self.grad.zero_()
        self.touched.clear()
        for p, o in zip(self.params, self.offsets):          # re-attach if something reset .grad
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def active_blocks(self) -> torch.Tensor:
        """uint8 mask, one byte per ALIGN elements of the flat buffer: 1 where the owning parameter
        was touched since the last zero_grad() (host tensor)."""
        m = torch.zeros(self.numel // self.ALIGN, dtype=torch.uint8)
        for i in self.touched:
            o, n = self.offsets[i], self.params[i].numel()
            m[o // self.ALIGN:(o + n + self.ALIGN - 1) // self.ALIGN] = 1
        return m

    def sync(self, force: bool = False) -> None:
        """Refresh the bf16 shadow with one kernel if any parameter changed since the last sync."""
        if K.compute_dtype() != torch.bfloat16:
            return
        if self.shadow16 is None:
            self.shadow16 = torch.empty(self.numel, dtype=torch.bfloat16, device=self.flat.device)
            for p, o in zip(self.params, self.offsets):
                self._views16[id(p)] = self.shadow16[o:o + p.numel()].view(p.shape)
            force = True
        # `p.data = view` keeps each parameter's own version counter, so sum them (in-place
        # optimizer / load_state_dict updates bump them; raw-pointer updates by dl_* kernels
        # that also write the shadow do not need a refresh)
        ver = self.flat._version + sum(p._version for p in self.params)
        if force or ver != self._synced:
            K.cast(self.flat, torch.bfloat16, out=self.shadow16)
            self._synced = ver

    def mark_synced(self) -> None:
        self._synced = self.flat._version + sum(p._version for p in self.params)

    def shadow_of(self, p: torch.Tensor) -> torch.Tensor:
        if K.compute_dtype() == torch.float32:
            return p.detach()
        self.sync()
        return self._views16[id(p)]


def fused_group(ps):
    """(weights, grad, shadow) views of the parameters `ps` concatenated along dim 0, when they were
    laid out as one group of a FlatParams store (else None).  `shadow` is in the compute dtype."""
    key = tuple(p.data_ptr() for p in ps)           # device addresses: robust to re-wrapped tensors
    ent = _GROUPS.get(key)
    if ent is None:
        return None
    fp = ent[0]()
    if fp is None:                                   # the store is gone (addresses may be reused)
        _GROUPS.pop(key, None)
        return None
    o, shape = ent[1], ent[2]
    if any(p.dtype != torch.float32 for p in ps) or fp.flat.data_ptr() + 4 * o != key[0]:
        return None
    n = 1
    for d in shape:
        n *= d
    w = fp.flat[o:o + n].view(shape)
    g = fp.grad[o:o + n].view(shape)
    if K.compute_dtype() == torch.bfloat16:
        fp.sync()
        sh = fp.shadow16[o:o + n].view(shape)
    else:
        sh = w
    return w, g, sh


class FlatAdamW:
    """AdamW (torch.optim.AdamW semantics) over a :class:`FlatParams` store: one kernel updates all
    parameters, their Adam moments and the bf16 shadow.  CUDA-graph capturable (the step counter
    lives on the device)."""

    def __init__(self, flat: FlatParams, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 per_param_steps: bool = False):
        """per_param_steps: keep one step count per parameter (torch's state['step']): needed when the
        set of parameters that receive a gradient changes between steps (trainer_step.TrainerStep); with
        a constant set (train.TrainStep) the single device counter is the same thing."""
        self.flat, self.lr, self.betas, self.eps, self.weight_decay = flat, lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.step_count = torch.zeros((), dtype=torch.int64, device=flat.flat.device)
        self.step_blocks = (torch.zeros(flat.numel // flat.ALIGN, dtype=torch.int32, device=flat.flat.device)
                            if per_param_steps else None)
        # device copy of FlatParams.active_blocks(); refreshed when the set of touched parameters
        # changes (never inside a CUDA-graph capture: TrainStep.capture refreshes it before capturing)
        self.active = torch.zeros(flat.numel // flat.ALIGN, dtype=torch.uint8, device=flat.flat.device)
        self._active_key = None

    def refresh_active(self) -> None:
        key = frozenset(self.flat.touched)
        if key != self._active_key:
            self.active.copy_(self.flat.active_blocks())
            self._active_key = key

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.flat.zero_grad()

    def step(self, grad_scale: float = 1.0, lo: int = 0, hi: Optional[int] = None, tick: bool = True) -> None:
        """One AdamW step over elements [lo, hi) of the flat buffer (default: all of it).  A step may be
        issued as several ranges; exactly one of them -- the LAST -- ticks the step counter (the earlier ones
        use count + 1 for their bias correction, dl_adamw_step).  lo / hi must be multiples of ALIGN."""
        from . import _lib as L
        f = self.flat
        hi = f.numel if hi is None else hi
        if lo % f.ALIGN or hi % f.ALIGN or not 0 <= lo <= hi <= f.numel:
            raise ValueError("AdamW range must lie on parameter-block boundaries")
        f.sync()
        if not torch.cuda.is_current_stream_capturing():
            self.refresh_active()
        bf = f.shadow16 is not None and K.compute_dtype() == torch.bfloat16
        sh = f.shadow16.data_ptr() + 2 * lo if bf else None
        blk = lo // f.ALIGN
        L.call("dl_adamw_step", f.flat.data_ptr() + 4 * lo, f.grad.data_ptr() + 4 * lo,
               self.exp_avg.data_ptr() + 4 * lo, self.exp_avg_sq.data_ptr() + 4 * lo, sh, hi - lo,
               self.step_count.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
               grad_scale, self.active.data_ptr() + blk,
               None if self.step_blocks is None else self.step_blocks.data_ptr() + 4 * blk, int(tick))
        f.mark_synced()       # raw-pointer update: versions unchanged, shadow already fresh
