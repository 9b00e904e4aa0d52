"""torch.autograd.Function wrappers: forward AND backward of every op are hand-written sm_100a
kernels (``kernels.py`` -> C ABI).  PyTorch only does the autograd bookkeeping.

Activations flow in the compute dtype (``kernels.compute_dtype()``: fp32 -> TF32 tensor cores,
bf16 -> bf16 tensor cores); parameters stay fp32 masters with compute-dtype shadows
(``params.shadow``); parameter gradients are produced in fp32.
"""
from __future__ import annotations

import contextlib
import itertools
import os
from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L
from . import kernels as K
from .params import flat_grad_of, fused_group, mark_touched, shadow

_seed_counter = itertools.count(0x5EED)
_seed_stream = 0


def set_seed_stream(rank: int) -> None:
    """Give every data-parallel rank its own dropout stream (folded into each launch seed)."""
    global _seed_stream
    _seed_stream = (int(rank) * 0xA24BAED4963EE407) & 0xFFFFFFFFFFFFFFFF


def next_seed() -> int:
    """A fresh dropout seed; masks are regenerated from it in backward."""
    return ((next(_seed_counter) * 0x9E3779B97F4A7C15) ^ _seed_stream) & 0xFFFFFFFFFFFFFFFF


# ---- branch streams -----------------------------------------------------------------------------------
# Independent branches of the forward (drug / protein extractors and their guided attentions; the
# protein and molecule streams of the paired PMMA blocks) are issued on two CUDA streams.  autograd runs
# every backward node on the stream of its forward, so the backward forks the same way.  Works eagerly
# and inside a CUDA-graph capture (the side stream joins the capture through its wait on the capturing
# stream).  DL_NO_BRANCH_STREAMS=1 turns it off (A/B measurements).
BRANCH_STREAMS = os.environ.get("DL_NO_BRANCH_STREAMS", "0") == "0"
_branch_streams = {}


def branch_stream(like: torch.Tensor, index: int = 0):
    """Side stream `index` of `like`'s device (None when branch streams are off or `like` is not on a GPU)."""
    if not BRANCH_STREAMS or not like.is_cuda:
        return None
    key = (like.device, index)
    s = _branch_streams.get(key)
    if s is None:
        s = _branch_streams[key] = torch.cuda.Stream(like.device)
    return s


def crosses(stream, *tensors):
    """Tensors allocated on one stream and read on another: tell the caching allocator, so their blocks
    are not handed out again on the allocating stream while the reader is still pending."""
    for t in tensors:
        if torch.is_tensor(t) and t.is_cuda:
            t.record_stream(stream)


_forward_only = False


@contextlib.contextmanager
def forward_only():
    """Forward-only scoring (druglamp_b200/infer.py): inside, the Functions skip the stores that
    only a backward pass reads (GELU / ReLU pre-activations).  A Function's forward cannot see the
    caller's grad mode (it always runs with grad disabled and ``needs_input_grad`` ignores
    ``torch.no_grad``), hence the explicit switch."""
    global _forward_only
    prev, _forward_only = _forward_only, True
    try:
        yield
    finally:
        _forward_only = prev


def _needs_backward(ctx) -> bool:
    return (not _forward_only) and any(ctx.needs_input_grad)


def _back(g: Optional[torch.Tensor], like_dtype: torch.dtype, shape=None):
    if g is None:
        return None
    if g.dtype != like_dtype:
        g = K.cast(g, like_dtype)
    return g if shape is None else g.reshape(shape)


def _as(g: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Incoming gradient, contiguous, in the dtype this Function's forward computed in (which may
    differ from the global compute dtype: kernels.local_compute_dtype)."""
    return K.cast(g if g.is_contiguous() else g.contiguous(), dtype)


def _grad_target(p):
    """The fp32 gradient buffer of a parameter that lives in a FlatParams store (None otherwise).
    Weight / bias gradients are then accumulated straight into it by the producing kernel
    (dl_gemm accumulate / dl_colsum accumulate) and the Function returns None for that input, which
    saves a temporary, a memset and an `at::add` launch per parameter."""
    if p is None or not torch.is_tensor(p):
        return None
    g = flat_grad_of(p)
    return g if (g is not None and g.is_contiguous()) else None


# ---- deferred parameter gradients -------------------------------------------------------------------
# Nothing in the backward reads a parameter gradient, yet the weight-gradient GEMMs (split-K over
# 16k-32k rows, a third of the step's GEMM time) sit in the middle of the dX chain.  Inside
# `deferred_weight_grads()` every kernel that ACCUMULATES INTO THE FLAT GRADIENT BUFFER is issued on a
# side stream that waits for its operands; the caller's stream only waits for that stream when the
# context closes (train.TrainStep: right after loss.backward()).  Outside the context nothing changes:
# a user who reads p.grad after backward() sees stream-ordered results as with any PyTorch module.
_wgrad_defer = None


class deferred_weight_grads:
    """Context: parameter-gradient kernels with a flat-buffer destination run on `stream` (created on
    first use); closing the context makes the current stream wait for them."""

    _streams = {}

    N_STREAMS = int(os.environ.get("DL_WGRAD_STREAMS", "3"))

    def __init__(self, enabled: bool = True):
        self.enabled = enabled and os.environ.get("DL_NO_WGRAD_STREAM", "0") == "0"
        self.streams = []       # created on first use; kernels go to them round robin (one serial queue
        self._next = 0          # of ~60 split-K GEMMs would otherwise drain long after the dX chain ends)
        self.keep = []          # operands stay referenced until the join: the autograd engine adds
        #                         gradients IN PLACE into a buffer it holds the last reference to

    def __enter__(self):
        global _wgrad_defer
        self._prev, _wgrad_defer = _wgrad_defer, (self if self.enabled else None)
        return self

    def join(self, stream=None):
        """`stream` (default: the current one) waits for every deferred kernel issued so far."""
        for s in self.streams:
            (stream or torch.cuda.current_stream()).wait_stream(s)

    def __exit__(self, *exc):
        global _wgrad_defer
        _wgrad_defer = self._prev
        self.join()
        self.keep.clear()
        return False

    @contextlib.contextmanager
    def on(self, *reads):
        cur = torch.cuda.current_stream()
        if not self.streams:
            dev = cur.device
            if dev not in self._streams:
                self._streams[dev] = [torch.cuda.Stream(dev) for _ in range(max(1, self.N_STREAMS))]
            self.streams = self._streams[dev]
        s = self.streams[self._next % len(self.streams)]
        self._next += 1
        s.wait_stream(cur)
        for t in reads:
            if t is not None:
                t.record_stream(s)
                self.keep.append(t)
        with torch.cuda.stream(s):
            yield


@contextlib.contextmanager
def _param_grad_stream(*reads):
    """Where the kernels inside may run: the deferred stream when a deferral is active."""
    if _wgrad_defer is None:
        yield
    else:
        with _wgrad_defer.on(*reads):
            yield


def _wgrad(p, a, b):
    """dW = a^T b (both stored [rows, features]); into p.grad when possible."""
    tgt = _grad_target(p)
    if tgt is not None and tgt.dim() == 2:
        with _param_grad_stream(a, b):
            K.mm(a, b, tgt, ta=True, tb=True, accumulate=True)
        return None
    return K.mm(a, b, ta=True, tb=True, out_dtype=torch.float32)


def _bgrad(p, g):
    tgt = _grad_target(p)
    if tgt is not None:
        with _param_grad_stream(g):
            K.colsum(g, tgt, accumulate=True)
        return None
    return K.colsum(g)


def _wbgrad(wp, bp, g, x):
    """(dW, db) of y = x W^T + b given g = dL/dy.  With a flat gradient store and bf16 operands this
    is ONE launch: the weight-gradient GEMM dW = g^T x also sums g's columns from the tiles it
    streams through shared memory (dl_gemm colsum_a), which is the bias gradient."""
    tw, tb = _grad_target(wp), _grad_target(bp)
    if tw is not None and tw.dim() == 2 and tb is not None and g.dtype == torch.bfloat16:
        with _param_grad_stream(g, x):
            K.mm(g, x, tw, ta=True, tb=True, accumulate=True, colsum_a=tb)
        return None, None
    return _wgrad(wp, g, x), (None if bp is None else _bgrad(bp, g))


class CastFn(Function):
    """dtype conversion with the dl_cast kernel (the gradient is converted back)."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.xdt = x.dtype
        return K.cast(x, dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return _back(g.contiguous(), ctx.xdt), None


# ================================================================================ Linear
def _align() -> int:
    """TMA needs 16-byte row strides: 8 bf16 or 4 fp32 elements."""
    return 8 if K.compute_dtype() == torch.bfloat16 else 4


def _pad_cols(t: torch.Tensor, n: int) -> torch.Tensor:
    return t if t.shape[-1] == n else torch.nn.functional.pad(t, (0, n - t.shape[-1]))


class LinearFn(Function):
    """y = dropout(act(x W^T + b)) + residual   (one GEMM, everything else in its epilogue).
    w is [out, in] (nn.Linear) or, with w_kn=True, [in, out] (GraphConv's layout).  Feature
    widths that are not a multiple of the TMA alignment (75, 385, 641, 1) are zero-padded."""

    @staticmethod
    def forward(ctx, x, w, b, act, residual, drop_p, seed, w_kn, keep_pad=False):
        """x may already carry the alignment padding of the weight's input width (zero columns, e.g.
        the 648-wide fill-bit concat for a 641-input layer): it is then used as is.  keep_pad
        returns the output with its alignment padding (zero columns) instead of slicing it off, so
        a chain of odd-width layers runs without padding or compaction copies."""
        xs = x.shape
        Kd = xs[-1]
        Kw = w.shape[0] if w_kn else w.shape[1]
        N = w.shape[1] if w_kn else w.shape[0]
        al = _align()
        Kp, Np = -(-Kw // al) * al, -(-N // al) * al
        if Kd != Kw and Kd != Kp:
            raise ValueError(f"input width {Kd} matches neither the weight's {Kw} nor its padded {Kp}")
        x2 = _pad_cols(K.to_compute(x).view(-1, Kd), Kp)
        wc = shadow(w)
        if Kp != Kw or Np != N:
            pad = (0, Np - N, 0, Kp - Kw) if w_kn else (0, Kp - Kw, 0, Np - N)
            wc = torch.nn.functional.pad(wc, pad)
        bias = None if b is None else _pad_cols(b.detach(), Np)
        out = torch.empty((x2.shape[0], Np), dtype=x2.dtype, device=x2.device)
        pre = torch.empty_like(out) if (act != K.ACT_NONE and _needs_backward(ctx)) else None
        r2 = None
        if residual is not None:
            r2 = _pad_cols(K.to_compute(residual).view(-1, residual.shape[-1]), Np)
        K.mm(x2, wc, out, tb=w_kn, bias=bias, act=act, pre=pre, res=r2, drop=(drop_p, seed))
        ctx.save_for_backward(x2, wc, pre)
        ctx.precise = L.FP32_PRECISE
        ctx.params = (w, b)
        No = Np if keep_pad else N
        ctx.meta = (act, drop_p, seed, xs, x.dtype, None if residual is None else residual.dtype,
                    b is not None, w_kn, Kd, Kw, N, No, None if residual is None else residual.shape)
        y = out if Np == No else out[:, :N]
        return y.reshape(*xs[:-1], No)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        with K.fp32_precise(ctx.precise):
            return LinearFn._backward(ctx, gy)

    @staticmethod
    def _backward(ctx, gy):
        x2, wc, pre = ctx.saved_tensors
        act, drop_p, seed, xs, xdt, rdt, has_b, w_kn, Kd, Kw, N, No, rshape = ctx.meta
        Kp = x2.shape[1]
        Np = wc.shape[1] if w_kn else wc.shape[0]
        gy2 = _as(gy, x2.dtype).view(-1, No)
        g = K.act_bwd(_pad_cols(gy2, Np), pre, act, (drop_p, seed))
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            dx = K.mm(g, wc, tb=not w_kn)
            dx = _back(dx if Kp == Kd else dx[:, :Kd], xdt, xs)
        wp, bp = ctx.params
        exact = Kp == Kw and Np == N
        if exact and not w_kn and has_b and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
            dw, db = _wbgrad(wp, bp, g, x2)
        else:
            if ctx.needs_input_grad[1]:
                if w_kn:
                    dw = _wgrad(wp, x2, g) if exact else K.mm(x2, g, ta=True, tb=True, out_dtype=torch.float32)[:Kw, :N]
                else:
                    dw = _wgrad(wp, g, x2) if exact else K.mm(g, x2, ta=True, tb=True, out_dtype=torch.float32)[:N, :Kw]
            if has_b and ctx.needs_input_grad[2]:
                db = _bgrad(bp, g) if exact else K.colsum(g)[:N]
        if rdt is not None and ctx.needs_input_grad[4]:
            gr = gy2 if rshape[-1] == No else gy2[:, :rshape[-1]]
            dres = _back(gr, rdt, rshape)
        return dx, dw, db, None, dres, None, None, None, None


class HeadLayerFn(Function):
    """y = BatchNorm1d(act(x W^T + b)) on at most K.SMALL_M rows, fp32: one layer of the decoder head
    (model/basic_model.py:205-213) as one launch forward (dl_small_linear) and two backward
    (dl_head_bn_act_bwd, then dX = g W by dl_small_linear); dW = g^T x is a dl_gemm that nothing
    downstream waits for.  bn = None: plain Linear (+ activation)."""

    @staticmethod
    def forward(ctx, x, w, b, act, bn_w, bn_b, running_mean, running_var, nbt, eps, momentum, training, has_bn):
        xs = x.shape
        x2 = x.reshape(-1, xs[-1]).float().contiguous()
        need = _needs_backward(ctx)
        bn = (None if bn_w is None else bn_w.detach(), None if bn_b is None else bn_b.detach(), running_mean,
              running_var, nbt, eps, momentum, training) if has_bn else None
        y, pre, mean, rstd = K.small_linear(x2, w.detach(), None if b is None else b.detach(), act=act,
                                            keep_pre=need and (has_bn or act != K.ACT_NONE), bn=bn)
        ctx.save_for_backward(x2, w, pre, mean, rstd, bn_w)
        ctx.params = (b, bn_b)
        ctx.meta = (act, bool(training), has_bn, xs, x.dtype)
        return y.view(*xs[:-1], w.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2, w, pre, mean, rstd, bn_w = ctx.saved_tensors
        b, bn_b = ctx.params
        act, training, has_bn, xs, xdt = ctx.meta
        N = w.shape[0]
        g = gy.reshape(-1, N).float().contiguous()
        dgam = dbet = db = None
        if pre is not None:
            tg, tb_, tbias = _grad_target(bn_w), _grad_target(bn_b), _grad_target(b)
            flat = (not has_bn or bn_w is None or (tg is not None and tb_ is not None)) and (b is None or tbias is not None)
            if flat:
                g, _, _, _ = K.head_bn_act_bwd(g, pre, None if bn_w is None else bn_w.detach(), mean, rstd, act,
                                               training, b is not None, acc_into=(tg, tb_, tbias))
            else:
                g, dgam, dbet, db = K.head_bn_act_bwd(g, pre, None if bn_w is None else bn_w.detach(), mean, rstd,
                                                      act, training, b is not None)
        elif b is not None:
            db = g.sum(0) if N <= K.SMALL_M else _bgrad(b, g)
        dx = None
        if ctx.needs_input_grad[0]:
            dx, _, _, _ = K.small_linear(g, w.detach(), None, w_kn=True)
            dx = _back(dx, xdt, xs)
        dw = None
        if ctx.needs_input_grad[1]:
            if N <= K.SMALL_M:       # (N, rows) x (rows, K): narrower than a TMA row -- the small-M kernel again
                dw = K.small_linear(g.t().contiguous(), x2, None, w_kn=True)[0]
            else:
                dw = _wgrad(w, g, x2)
        return dx, dw, db, None, dgam, dbet, None, None, None, None, None, None, None


def head_layer(x, fc: torch.nn.Linear, act=K.ACT_NONE, bn: Optional[torch.nn.BatchNorm1d] = None):
    """fc -> act -> bn (optional) on <= K.SMALL_M rows through the small-M kernels."""
    if bn is None:
        return HeadLayerFn.apply(x, fc.weight, fc.bias, act, None, None, None, None, None, 0.0, 0.0, False, False)
    training = bn.training or bn.running_mean is None
    momentum = 0.1 if bn.momentum is None else bn.momentum
    update = training and not _bn_frozen
    keep = update or not training
    return HeadLayerFn.apply(x, fc.weight, fc.bias, act, bn.weight, bn.bias, bn.running_mean if keep else None,
                             bn.running_var if keep else None, bn.num_batches_tracked if update else None,
                             bn.eps, momentum, training, True)


def linear(x, w, b=None, act=K.ACT_NONE, residual=None, drop_p=0.0, seed=0, keep_pad=False):
    return LinearFn.apply(x, w, b, act, residual, drop_p, seed, False, keep_pad)


# ================================================================================ FFN
# DL_NO_FUSED_FFN=1: the two-GEMM path (A/B measurements; also what fp32 mode and other widths use)
FUSED_FFN = os.environ.get("DL_NO_FUSED_FFN", "0") == "0"


def _al(t, nbytes) -> bool:
    return t is not None and t.data_ptr() % nbytes == 0


class FFNFn(Function):
    """y = dropout(fc2(dropout(gelu(fc1(x))))) + residual  (PMMA Mlp, model/PMMA/mlp.py:44-50,
    with the block's residual add, model/PMMA/block.py:45-47).  bf16 rows of width 256: ONE launch
    forward (dl_ffn_fwd: both GEMMs chained on chip, the hidden activation and its derivative stored
    once for the backward) and ONE for (dpre, dX) backward (dl_ffn_bwd).  Otherwise two GEMMs each
    way with gelu' and the first dropout mask fused into the dX GEMM epilogue."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual, p, seed1, seed2):
        xs = x.shape
        x2 = K.to_compute(x).view(-1, xs[-1])
        w1c, w2c = shadow(w1), shadow(w2)
        M, Dh = x2.shape[0], w1.shape[0]
        r2 = K.to_compute(residual).view(-1, w2.shape[0]) if residual is not None else None
        need = _needs_backward(ctx)
        fused = (FUSED_FFN and K.ffn_supported(x2, w1c, w2c) and _al(b1, 16) and _al(b2, 16)
                 and b1.dtype == torch.float32 and b2.dtype == torch.float32
                 and (r2 is None or (_al(r2, 32) and r2.stride(0) % 16 == 0 and r2.dtype == x2.dtype)))
        if fused:
            y, hd, pre1 = K.ffn_fwd(x2, w1c, b1.detach(), w2c, b2.detach(), r2, (p, seed1, seed2), keep=need)
        else:
            hd = torch.empty((M, Dh), dtype=x2.dtype, device=x2.device)
            # forward-only scoring (no_grad): the pre-activation is never read, skip its store
            pre1 = torch.empty_like(hd) if need else None
            # pre1 receives d hd / d pre = gelu'(pre) * dropout factor (pre_mode 1): the backward multiplies
            K.mm(x2, w1c, hd, bias=b1, act=K.ACT_GELU, pre=pre1, drop=(p, seed1), pre_mode=1)
            y = K.mm(hd, w2c, bias=b2, res=r2, drop=(p, seed2))
        ctx.save_for_backward(x2, w1, w2, pre1, hd)
        ctx.biases = (b1, b2)
        ctx.meta = (p, seed1, seed2, xs, x.dtype, None if residual is None else residual.dtype, fused)
        return y.view(*xs[:-1], w2.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2, w1, w2, pre1, hd = ctx.saved_tensors
        p, seed1, seed2, xs, xdt, rdt, fused = ctx.meta
        gy2 = K.to_compute(gy).view(-1, w2.shape[0])
        g2 = K.act_bwd(gy2, None, K.ACT_NONE, (p, seed2))
        b1p, b2p = ctx.biases
        dw2, db2 = _wbgrad(w2, b2p, g2, hd)
        dx = None
        if fused and ctx.needs_input_grad[0] and _al(g2, 32):
            dpre1, dxc = K.ffn_bwd(g2, shadow(w1), shadow(w2), pre1)
            dx = _back(dxc, xdt, xs)
        else:
            dpre1 = K.mm(g2, shadow(w2), tb=True, mul_aux=pre1, mul_mode=K.MUL_VALUE)
        dw1, db1 = _wbgrad(w1, b1p, dpre1, x2)
        if dx is None and ctx.needs_input_grad[0]:
            dx = _back(K.mm(dpre1, shadow(w1), tb=True), xdt, xs)
        dres = _back(gy2, rdt, gy.shape) if rdt is not None else None
        return dx, dw1, db1, dw2, db2, dres, None, None, None


def ffn(x, w1, b1, w2, b2, residual=None, p=0.0):
    s1, s2 = (next_seed(), next_seed()) if p > 0 else (0, 0)
    return FFNFn.apply(x, w1, b1, w2, b2, residual, p, s1, s2)


# ================================================================================ LayerNorm
class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        xc = K.to_compute(x)
        y, mean, rstd = K.layernorm_fwd(xc, gamma.detach(), beta.detach(), eps)
        ctx.save_for_backward(xc, gamma, beta, mean, rstd)
        ctx.xdt = x.dtype
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xc, gamma, beta, mean, rstd = ctx.saved_tensors
        tg, tb = _grad_target(gamma), _grad_target(beta)
        if tg is not None and tb is not None:       # straight into the flat gradient buffer
            dx, _, _ = K.layernorm_bwd(K.to_compute(gy), xc, gamma.detach(), mean, rstd, acc_into=(tg, tb))
            return _back(dx, ctx.xdt), None, None, None
        dx, dg, db = K.layernorm_bwd(K.to_compute(gy), xc, gamma.detach(), mean, rstd)
        return _back(dx, ctx.xdt), dg, db, None


def layer_norm(x, gamma, beta, eps=1e-5):
    return LayerNormFn.apply(x, gamma, beta, eps)


class LayerNormResFn(Function):
    """(LayerNorm(x), x): the pre-norm residual pattern ``x + f(LN(x))`` (model/PMMA/block.py:33-47).
    Handing the skip connection out of the same Function lets the backward add its gradient to the
    LayerNorm's input gradient inside dl_layernorm_bwd instead of in autograd's separate add pass."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        xc = K.to_compute(x)
        y, mean, rstd = K.layernorm_fwd(xc, gamma.detach(), beta.detach(), eps)
        ctx.save_for_backward(xc, gamma, beta, mean, rstd)
        ctx.xdt = x.dtype
        return y, x.view_as(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gres):
        xc, gamma, beta, mean, rstd = ctx.saved_tensors
        add = None
        if gres is not None:
            add = K.to_compute(gres).view(xc.shape)
        tg, tb = _grad_target(gamma), _grad_target(beta)
        if tg is not None and tb is not None:
            dx, _, _ = K.layernorm_bwd(K.to_compute(gy), xc, gamma.detach(), mean, rstd, acc_into=(tg, tb),
                                       dx_add=add)
            return _back(dx, ctx.xdt), None, None, None
        dx, dg, db = K.layernorm_bwd(K.to_compute(gy), xc, gamma.detach(), mean, rstd, dx_add=add)
        return _back(dx, ctx.xdt), dg, db, None


def layer_norm_res(x, gamma, beta, eps=1e-5):
    """-> (LayerNorm(x), x) with the skip connection's gradient fused into the LayerNorm backward."""
    return LayerNormResFn.apply(x, gamma, beta, eps)


# ================================================================================ attention core
def _score_buf(B, H, S2, Lq, Lk, like):
    """(B,H,S2,Lq,Lk) score / probability map whose rows are padded to the 16-byte stride TMA
    needs (a key length such as 290 is legal for GuidedCrossAttention); the pad columns are never
    read: every consumer addresses the map with extent Lk."""
    q = 16 // like.element_size()
    Lkp = (Lk + q - 1) // q * q
    full = torch.empty((B, H, S2, Lq, Lkp), dtype=like.dtype, device=like.device)
    return full if Lkp == Lk else full[..., :Lk]


def _attn_fwd(q, k, v, H, scale, keep_raw, out=None):
    """q (S2,B,Lq,HD), k, v (B,Lk,HD): unit inner stride, every other stride free (they may be
    column slices of one fused QKV projection buffer).
    Returns O (B,Lq,S2*HD), the tensor the backward needs besides O (the fused kernel's row
    log-sum-exp (S2,B,H,Lq) fp32, or the probability map P (B,H,S2,Lq,Lk) of the unfused path) and
    the raw scaled logits (B,H,S2,Lq,Lk) (or None)."""
    S2, B, Lq, HD = q.shape
    Lk = k.shape[1]
    d = HD // H
    if K.attn_supported(q, k, H) and (S2 == 1 or not keep_raw):
        # ONE tcgen05 kernel: scores, softmax and p.v never leave the SM; only the row
        # log-sum-exp is kept for the backward (which recomputes the probabilities)
        O, lse, raw = K.attn_fwd(q, k, v, H, scale, want_raw=keep_raw, out=out)
        return O, lse, (None if raw is None else raw.unsqueeze(2))
    if out is not None:
        raise ValueError("a preallocated output needs the fused attention kernel")
    S = _score_buf(B, H, S2, Lq, Lk, q)
    ldp = S.stride(3)
    s_lay = (S.stride(1), S.stride(2), S.stride(0))       # batch order (head, set, pair)
    L.gemm(q, k, S, M=Lq, N=Lk, K=d, lda=q.stride(2), ldb=k.stride(1), ldc=ldp, batch=(H, S2, B),
           sa=(d, q.stride(0) if S2 > 1 else 0, q.stride(1)), sb=(d, 0, k.stride(0)), sc=s_lay,
           alpha=scale)
    raw = None
    if keep_raw:
        raw = S
        P = K.softmax_fwd(S, out=_score_buf(B, H, S2, Lq, Lk, q))
    else:
        P = K.softmax_fwd(S)
    O = torch.empty((B, Lq, S2 * HD), dtype=q.dtype, device=q.device)
    L.gemm(P, v, O, M=Lq, N=d, K=Lk, lda=ldp, ldb=v.stride(1), ldc=S2 * HD, trans_b=True,
           batch=(H, S2, B), sa=s_lay, sb=(d, 0, v.stride(0)), sc=(d, HD, Lq * S2 * HD))
    return O, P, raw


def _attn_bwd(dO, q, k, v, P, O, H, scale, dq_out=None, dk_out=None, dv_out=None, dq_accumulate=False):
    """Gradients of _attn_fwd (P = its second result, O its first).  d*_out may be preallocated
    (strided) destinations; with dq_accumulate the query gradient is added to dq_out (two
    attentions sharing one query set)."""
    S2, B, Lq, HD = q.shape
    Lk = k.shape[1]
    d = HD // H
    if P.dtype == torch.float32 and P.dim() == 4 and q.dtype == torch.bfloat16:     # fused: P is the lse
        dV = dv_out if dv_out is not None else torch.empty((B, Lk, HD), dtype=q.dtype, device=q.device)
        dK = dk_out if dk_out is not None else torch.empty((B, Lk, HD), dtype=q.dtype, device=q.device)
        dQ = dq_out if dq_out is not None else torch.empty((S2, B, Lq, HD), dtype=q.dtype, device=q.device)
        K.attn_bwd(dO, q, k, v, O, P, H, scale, dQ, dK, dV, dq_accumulate=dq_accumulate)
        return dQ, dK, dV
    ldp = P.stride(3)
    s_lay = (P.stride(1), P.stride(2), P.stride(0))
    dV = dv_out if dv_out is not None else torch.empty((B, Lk, HD), dtype=q.dtype, device=q.device)
    dK = dk_out if dk_out is not None else torch.empty((B, Lk, HD), dtype=q.dtype, device=q.device)
    dQ = dq_out if dq_out is not None else torch.empty_like(q)
    # dV = sum_set P_set^T dO_set: the query-set dimension is reduced inside one GEMM (kred = 2)
    L.gemm(P, dO, dV, M=Lk, N=d, K=Lq, lda=ldp, ldb=S2 * HD, ldc=dV.stride(1), trans_a=True,
           trans_b=True, batch=(H, B, S2), sa=(s_lay[0], s_lay[2], s_lay[1]), sb=(d, Lq * S2 * HD, HD),
           sc=(d, dV.stride(0), 0), kred=2)
    # dP = dO V^T, then dS = scale * P * (dP - rowsum(P dP)) in place
    dP = _score_buf(B, H, S2, Lq, Lk, P)
    L.gemm(dO, v, dP, M=Lq, N=Lk, K=d, lda=S2 * HD, ldb=v.stride(1), ldc=ldp, batch=(H, S2, B),
           sa=(d, HD, Lq * S2 * HD), sb=(d, 0, v.stride(0)), sc=s_lay)
    dS = K.softmax_bwd(P, dP, scale)
    # dQ = dS K
    L.gemm(dS, k, dQ, M=Lq, N=d, K=Lk, lda=ldp, ldb=k.stride(1), ldc=dQ.stride(2), trans_b=True,
           batch=(H, S2, B), sa=s_lay, sb=(d, 0, k.stride(0)),
           sc=(d, dQ.stride(0) if S2 > 1 else 0, dQ.stride(1)), residual=dQ if dq_accumulate else None)
    # dK = sum_set dS_set^T Q_set, likewise
    L.gemm(dS, q, dK, M=Lk, N=d, K=Lq, lda=ldp, ldb=q.stride(2), ldc=dK.stride(1), trans_a=True,
           trans_b=True, batch=(H, B, S2), sa=(s_lay[0], s_lay[2], s_lay[1]),
           sb=(d, q.stride(1), q.stride(0) if S2 > 1 else Lq * q.stride(2)),
           sc=(d, dK.stride(0), 0), kred=2)
    return dQ, dK, dV


# ---- fused query / key / value projections -----------------------------------------------------
# The three projections of one input share a GEMM when their weights sit back to back in the flat
# parameter buffer (params.fused_group): one launch forward, one for dX (which also sums the three
# input gradients autograd would otherwise add), one for dW with the bias gradients riding on it.
# Without a flat store the same buffers are filled by three launches each.
def _qkv_fwd(xc2, ws, bs, out2):
    """out2 [rows, 3E] = xc2 [rows, D] @ cat(ws)^T + cat(bs)."""
    E = ws[0].shape[0]
    fw, fb = fused_group(ws), fused_group(bs)
    if fw is not None and fb is not None:
        K.mm(xc2, fw[2], out2, bias=fb[0])
    else:
        for i, (w, b) in enumerate(zip(ws, bs)):
            K.mm(xc2, shadow(w), out2[:, i * E:(i + 1) * E], bias=b.detach())


def _qkv_bwd(g2, xc2, ws, bs, need_dx=True):
    """g2 [rows, 3E] -> dx [rows, D]; parameter gradients accumulated in place when possible.
    Returns (dx, [dw_i], [db_i]) with None for everything that was accumulated."""
    E = ws[0].shape[0]
    fw, fb = fused_group(ws), fused_group(bs)
    dx = None
    if fw is not None and fb is not None and _grad_target(ws[0]) is not None:     # (None: in-place gradients off)
        for t in tuple(ws) + tuple(bs):
            mark_touched(t)
        if need_dx:
            dx = K.mm(g2, fw[2], tb=True)
        with _param_grad_stream(g2, xc2):
            if g2.dtype == torch.bfloat16:
                K.mm(g2, xc2, fw[1], ta=True, tb=True, accumulate=True, colsum_a=fb[1])
            else:
                K.mm(g2, xc2, fw[1], ta=True, tb=True, accumulate=True)
                K.colsum(g2, fb[1], accumulate=True)
        return dx, [None] * 3, [None] * 3
    dws, dbs = [], []
    for i, (w, b) in enumerate(zip(ws, bs)):
        gi = g2[:, i * E:(i + 1) * E]
        if need_dx:
            dx = K.mm(gi, shadow(w), tb=True) if dx is None else K.mm(gi, shadow(w), dx, tb=True, res=dx)
        dw, db = _wbgrad(w, b, gi, xc2)
        dws.append(dw)
        dbs.append(db)
    return dx, dws, dbs


class QKVProjFn(Function):
    """QKV (B, L, 3E) = [query(x), key(x), value(x)] in one buffer (model/PMMA/attention.py:109-111)."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv):
        xc = K.to_compute(x)
        D, E = xc.shape[-1], wq.shape[0]
        out = torch.empty(xc.shape[:-1] + (3 * E,), dtype=xc.dtype, device=xc.device)
        _qkv_fwd(xc.view(-1, D), (wq, wk, wv), (bq, bk, bv), out.view(-1, 3 * E))
        ctx.save_for_backward(xc, wq, wk, wv, bq, bk, bv)
        ctx.xdt = x.dtype
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        xc, wq, wk, wv, bq, bk, bv = ctx.saved_tensors
        D, E = xc.shape[-1], wq.shape[0]
        g2 = K.to_compute(g).view(-1, 3 * E)
        dx, dws, dbs = _qkv_bwd(g2, xc.view(-1, D), (wq, wk, wv), (bq, bk, bv), ctx.needs_input_grad[0])
        dx = None if dx is None else _back(dx, ctx.xdt, xc.shape)
        return dx, dws[0], dbs[0], dws[1], dbs[1], dws[2], dbs[2]


class PairedQKVFn(Function):
    """QKV (2, B, L, 3E): slab 0 = [query, key, value](prot), slab 1 = [query_mol, key_mol,
    value_mol](mol) (model/PMMA/attention.py:91-98), so the paired attention addresses both query
    sets and both K/V pairs of one buffer through strides."""

    @staticmethod
    def forward(ctx, xp, xm, *params):
        ap, am = K.to_compute(xp), K.to_compute(xm)
        wp, bp = params[0:6:2], params[1:6:2]
        wm, bm = params[6:12:2], params[7:12:2]
        D, E = ap.shape[-1], wp[0].shape[0]
        out = torch.empty((2,) + tuple(ap.shape[:-1]) + (3 * E,), dtype=ap.dtype, device=ap.device)
        side = branch_stream(ap)
        if side is None:
            _qkv_fwd(ap.view(-1, D), wp, bp, out[0].view(-1, 3 * E))
            _qkv_fwd(am.view(-1, D), wm, bm, out[1].view(-1, 3 * E))
        else:                                   # the molecule slab on the side stream, joined before the core
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                _qkv_fwd(am.view(-1, D), wm, bm, out[1].view(-1, 3 * E))
            _qkv_fwd(ap.view(-1, D), wp, bp, out[0].view(-1, 3 * E))
            main.wait_stream(side)
            crosses(side, out)
        ctx.save_for_backward(ap, am, *params)
        ctx.dts = (xp.dtype, xm.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        ap, am, *params = ctx.saved_tensors
        wp, bp = params[0:6:2], params[1:6:2]
        wm, bm = params[6:12:2], params[7:12:2]
        D, E = ap.shape[-1], wp[0].shape[0]
        gc = K.to_compute(g)
        side = branch_stream(gc)
        if side is None:
            dxp, dwp, dbp = _qkv_bwd(gc[0].view(-1, 3 * E), ap.view(-1, D), wp, bp, ctx.needs_input_grad[0])
            dxm, dwm, dbm = _qkv_bwd(gc[1].view(-1, 3 * E), am.view(-1, D), wm, bm, ctx.needs_input_grad[1])
        else:
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                crosses(side, gc)
                dxm, dwm, dbm = _qkv_bwd(gc[1].view(-1, 3 * E), am.view(-1, D), wm, bm, ctx.needs_input_grad[1])
            dxp, dwp, dbp = _qkv_bwd(gc[0].view(-1, 3 * E), ap.view(-1, D), wp, bp, ctx.needs_input_grad[0])
            main.wait_stream(side)          # autograd hands both results on from this node's stream
            crosses(main, dxm)
        grads = []
        for dw, db in zip(dwp, dbp):
            grads += [dw, db]
        for dw, db in zip(dwm, dbm):
            grads += [dw, db]
        return (None if dxp is None else _back(dxp, ctx.dts[0], ap.shape),
                None if dxm is None else _back(dxm, ctx.dts[1], am.shape), *grads)


class SelfAttnCoreFn(Function):
    """softmax(q k^T scale) v on a fused (B, L, 3E) QKV buffer -> (B, L, E); the backward writes
    dq / dk / dv straight into the three column blocks of one (B, L, 3E) gradient."""

    @staticmethod
    def forward(ctx, qkv, H, scale):
        c = K.to_compute(qkv)
        E = c.shape[-1] // 3
        O, P, _ = _attn_fwd(c[None, :, :, :E], c[:, :, E:2 * E], c[:, :, 2 * E:], H, scale, False)
        ctx.save_for_backward(c, P, O)
        ctx.meta = (H, scale, qkv.dtype)
        return O

    @staticmethod
    @once_differentiable
    def backward(ctx, gO):
        c, P, O = ctx.saved_tensors
        H, scale, dt = ctx.meta
        E = c.shape[-1] // 3
        d = torch.empty_like(c)
        _attn_bwd(K.to_compute(gO), c[None, :, :, :E], c[:, :, E:2 * E], c[:, :, 2 * E:], P, O, H, scale,
                  dq_out=d[None, :, :, :E], dk_out=d[:, :, E:2 * E], dv_out=d[:, :, 2 * E:])
        return _back(d, dt), None, None


class PairedAttnCoreFn(Function):
    """The four attention maps of the paired block (model/PMMA/attention.py:44-88) on a fused
    (2, B, L, 3E) QKV buffer: both query sets against the protein K/V -> op (B, L, 2E) =
    cat(attn, attn_p), and against the molecule K/V -> om (B, L, 2E) = (attn_p', attn') in swapped
    order.  The backward fills one (2, B, L, 3E) gradient; the query gradient of the second
    attention is accumulated onto the first in the GEMM epilogue."""

    @staticmethod
    def forward(ctx, qkv, H, scale):
        c = K.to_compute(qkv)
        E = c.shape[-1] // 3
        Q = c[:, :, :, :E]
        op, Pp, _ = _attn_fwd(Q, c[0, :, :, E:2 * E], c[0, :, :, 2 * E:], H, scale, False)
        om, Pm, _ = _attn_fwd(Q, c[1, :, :, E:2 * E], c[1, :, :, 2 * E:], H, scale, False)
        ctx.save_for_backward(c, Pp, Pm, op, om)
        ctx.meta = (H, scale, qkv.dtype)
        return op, om

    @staticmethod
    @once_differentiable
    def backward(ctx, gop, gom):
        c, Pp, Pm, op, om = ctx.saved_tensors
        H, scale, dt = ctx.meta
        E = c.shape[-1] // 3
        Q = c[:, :, :, :E]
        d = torch.empty_like(c)
        dQ = d[:, :, :, :E]
        _attn_bwd(K.to_compute(gop), Q, c[0, :, :, E:2 * E], c[0, :, :, 2 * E:], Pp, op, H, scale,
                  dq_out=dQ, dk_out=d[0, :, :, E:2 * E], dv_out=d[0, :, :, 2 * E:])
        _attn_bwd(K.to_compute(gom), Q, c[1, :, :, E:2 * E], c[1, :, :, 2 * E:], Pm, om, H, scale,
                  dq_out=dQ, dk_out=d[1, :, :, E:2 * E], dv_out=d[1, :, :, 2 * E:], dq_accumulate=True)
        return _back(d, dt), None, None


class FcCatFn(Function):
    """y = cat(A_first, A_second) W^T + b where the input holds the two halves as columns
    [0,D) and [D,2D).  swap=True means the buffer holds (A_second, A_first): the weight's column
    halves are addressed crosswise instead of copying (model/PMMA/attention.py:81)."""

    @staticmethod
    def forward(ctx, o, w, b, swap):
        oc = K.to_compute(o)
        D2 = oc.shape[-1]
        D = D2 // 2
        o2 = oc.view(-1, D2)
        wc = shadow(w)
        N = w.shape[0]
        if not swap:
            y = K.mm(o2, wc, bias=b)
        else:
            y = K.mm(o2[:, :D], wc[:, D:], bias=b)
            K.mm(o2[:, D:], wc[:, :D], y, res=y)
        ctx.save_for_backward(o2, w)
        ctx.bias = b
        ctx.meta = (swap, o.dtype, o.shape)
        return y.view(*o.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        o2, w = ctx.saved_tensors
        swap, odt, oshape = ctx.meta
        N, D2 = w.shape
        D = D2 // 2
        g = K.to_compute(gy).view(-1, N)
        wc = shadow(w)
        if not swap:
            do = K.mm(g, wc, tb=True)
            dw, db = _wbgrad(w, ctx.bias, g, o2)
        else:
            do = torch.empty_like(o2)
            K.mm(g, wc[:, D:], do[:, :D], tb=True)
            K.mm(g, wc[:, :D], do[:, D:], tb=True)
            tw, tb_ = _grad_target(w), _grad_target(ctx.bias)
            if tw is not None and tb_ is not None and g.dtype == torch.bfloat16:
                # crosswise halves straight into the flat gradient; the first GEMM also sums g's columns
                with _param_grad_stream(g, o2):
                    K.mm(g, o2[:, :D], tw[:, D:], ta=True, tb=True, accumulate=True, colsum_a=tb_)
                    K.mm(g, o2[:, D:], tw[:, :D], ta=True, tb=True, accumulate=True)
                dw = db = None
            else:
                dw = torch.empty((N, D2), dtype=torch.float32, device=g.device)
                K.mm(g, o2[:, :D], dw[:, D:], ta=True, tb=True)
                K.mm(g, o2[:, D:], dw[:, :D], ta=True, tb=True)
                db = _bgrad(ctx.bias, g)
        return _back(do, odt, oshape), dw, db, None


# ================================================================================ PGCA
def _rows(t: torch.Tensor, seq_first: bool) -> torch.Tensor:
    """A (L, N, E) tensor as the [L*N, E] row matrix of the chosen row order, in the compute dtype:
    seq_first keeps the reference's own (l, n) order (no copy for a contiguous input), otherwise the
    rows run (n, l) (no copy when the input is a permuted view of a batch-first tensor)."""
    Lr, Nr, E = t.shape
    return K.to_compute(t if seq_first else t.transpose(0, 1)).view(Lr * Nr, E)


def _blc(rows: torch.Tensor, Lr: int, Nr: int, seq_first: bool) -> torch.Tensor:
    """The (N, L, C) view of a row matrix built by _rows (strided when seq_first)."""
    C_ = rows.shape[1]
    return rows.view(Lr, Nr, C_).permute(1, 0, 2) if seq_first else rows.view(Nr, Lr, C_)


class PGCAFn(Function):
    """GuidedCrossAttention forward/backward (model/PGCA/guided_cross_attention_model.py:124-329,
    the enc-dec in-proj branch :138-162): in-proj GEMMs, scaled q k^T with the RAW logits kept
    (:307,:319-320), softmax, p v, out-proj.  Inputs are sequence-first (L,N,E) like the
    reference; the torch.equal host sync of :138 does not exist here.

    The projections are row-order agnostic and the attention kernel addresses its operands through
    strides, so a contiguous sequence-first input (what the reference API hands over, e.g. the
    1200 x 290 long-sequence configuration) is processed in place -- no (L,N,E) <-> (N,L,E) copies."""

    @staticmethod
    def forward(ctx, query, key, value, in_w, in_b, out_w, out_b, H):
        Lq, Bn, E = query.shape
        Sk = key.shape[0]
        shared = (key.data_ptr() == value.data_ptr() and key.stride() == value.stride()
                  and key.shape == value.shape)
        fused_ok = K.compute_dtype() == torch.bfloat16 and (E // H) in (64, 128) and Sk <= K.ATTN_MAX_KEYS
        sf = bool(fused_ok and query.is_contiguous() and key.is_contiguous() and value.is_contiguous()
                  and not query.transpose(0, 1).is_contiguous())
        q2, k2 = _rows(query, sf), _rows(key, sf)
        v2 = k2 if shared else _rows(value, sf)
        wi = shadow(in_w)
        ib = in_b.detach()
        Q2 = K.mm(q2, wi[:E], bias=ib[:E])
        KV2 = torch.empty((k2.shape[0], 2 * E), dtype=q2.dtype, device=q2.device)
        if shared:
            K.mm(k2, wi[E:], KV2, bias=ib[E:])
        else:
            K.mm(k2, wi[E:2 * E], KV2[:, :E], bias=ib[E:2 * E])
            K.mm(v2, wi[2 * E:], KV2[:, E:], bias=ib[2 * E:])
        Qp, KV = _blc(Q2, Lq, Bn, sf)[None], _blc(KV2, Sk, Bn, sf)
        scale = float(E // H) ** -0.5
        O2 = torch.empty((q2.shape[0], E), dtype=q2.dtype, device=q2.device) if sf else None
        O, P, raw = _attn_fwd(Qp, KV[:, :, :E], KV[:, :, E:], H, scale, True,
                              out=None if O2 is None else _blc(O2, Lq, Bn, sf))
        if O2 is None:
            O2 = O.view(-1, E)
        out2 = K.mm(O2, shadow(out_w), bias=out_b.detach())
        ctx.save_for_backward(q2, k2, v2, in_w, out_w, Q2, KV2, P, O2, in_b, out_b)
        ctx.meta = (H, scale, shared, sf, query.dtype, key.dtype, value.dtype, Lq, Sk, Bn)
        raw = raw[:, :, 0]                                             # (N, H, L, S)
        if not raw.is_contiguous():
            raw = raw.contiguous()                                     # padded rows (S % 8 != 0)
        ctx.mark_non_differentiable(raw)
        return (out2.view(Lq, Bn, E) if sf else out2.view(Bn, Lq, E).transpose(0, 1)), raw

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, _graw):
        q2, k2, v2, in_w, out_w, Q2, KV2, P, O2, in_b, out_b = ctx.saved_tensors
        H, scale, shared, sf, qdt, kdt, vdt, Lq, Sk, Bn = ctx.meta
        E = q2.shape[1]
        g = _rows(gout, sf)                                            # [L*N, E] in the forward's row order
        d_out_w, d_out_b = _wbgrad(out_w, out_b, g, O2)
        dO2 = K.mm(g, shadow(out_w), tb=True)
        dQ2, dKV2 = torch.empty_like(Q2), torch.empty_like(KV2)
        KV, dKV = _blc(KV2, Sk, Bn, sf), _blc(dKV2, Sk, Bn, sf)
        _attn_bwd(_blc(dO2, Lq, Bn, sf), _blc(Q2, Lq, Bn, sf)[None], KV[:, :, :E], KV[:, :, E:], P,
                  _blc(O2, Lq, Bn, sf), H, scale, dq_out=_blc(dQ2, Lq, Bn, sf)[None],
                  dk_out=dKV[:, :, :E], dv_out=dKV[:, :, E:])
        wi = shadow(in_w)

        def back(rows, Lr, dt):                                        # row matrix -> (L, N, E) gradient
            t = _back(rows, dt)
            return t.view(Lr, Bn, E) if sf else t.view(Bn, Lr, E).transpose(0, 1)
        # the packed in-proj gradients go straight into the flat gradient buffer when there is one
        tw, tb_ = _grad_target(in_w), _grad_target(in_b)
        acc = tw is not None and tb_ is not None
        d_in_w = tw if acc else torch.empty((3 * E, E), dtype=torch.float32, device=g.device)
        d_in_b = tb_ if acc else torch.empty(3 * E, dtype=torch.float32, device=g.device)
        fuse = acc and dQ2.dtype == torch.bfloat16      # bias gradients summed inside the dW GEMMs
        with (_param_grad_stream(dQ2, q2, dKV2, k2, v2) if acc else contextlib.nullcontext()):
            K.mm(dQ2, q2, d_in_w[:E], ta=True, tb=True, accumulate=acc, colsum_a=d_in_b[:E] if fuse else None)
            if not fuse:
                K.colsum(dQ2, d_in_b[:E], accumulate=acc)
            if not (fuse and shared):
                K.colsum(dKV2, d_in_b[E:], accumulate=acc)
            if shared:
                K.mm(dKV2, k2, d_in_w[E:], ta=True, tb=True, accumulate=acc, colsum_a=d_in_b[E:] if fuse else None)
            else:
                K.mm(dKV2[:, :E], k2, d_in_w[E:2 * E], ta=True, tb=True, accumulate=acc)
                K.mm(dKV2[:, E:], v2, d_in_w[2 * E:], ta=True, tb=True, accumulate=acc)
        dquery = back(K.mm(dQ2, wi[:E], tb=True), Lq, qdt)
        if shared:
            dkey = back(K.mm(dKV2, wi[E:], tb=True), Sk, kdt)
            dvalue = None
        else:
            dkey = back(K.mm(dKV2[:, :E], wi[E:2 * E], tb=True), Sk, kdt)
            dvalue = back(K.mm(dKV2[:, E:], wi[2 * E:], tb=True), Sk, vdt)
        if acc:
            d_in_w = d_in_b = None
        return dquery, dkey, dvalue, d_in_w, d_in_b, d_out_w, d_out_b, None


# ================================================================================ MHLA
class MHLAFn(Function):
    """MultiHeadLinearAttention (model/PMMA/encoder.py:127-140): logits = lin2(gelu(lin1(v))),
    softmax over the sequence, gating through the reinterpreting view.  With gamma/beta the
    residual add and LayerNorm of model/DrugLAMP.py:63-71 are fused in: y = LN(v + gate(v))."""

    @staticmethod
    def forward(ctx, v, w1, b1, w2, b2, gamma, beta, eps):
        vc = K.to_compute(v)
        Bn, Lr, E = vc.shape
        v2 = vc.view(-1, E)
        D, Hh = w1.shape[0], w2.shape[0]
        h = torch.empty((v2.shape[0], D), dtype=vc.dtype, device=vc.device)
        pre1 = torch.empty_like(h) if _needs_backward(ctx) else None
        K.mm(v2, shadow(w1), h, bias=b1, act=K.ACT_GELU, pre=pre1, pre_mode=1)      # pre1 = gelu'(lin1(v))
        logits = K.mm(h, shadow(w2), bias=b2).view(Bn, Lr, Hh)
        g_ = None if gamma is None else gamma.detach()
        b_ = None if beta is None else beta.detach()
        y, p, mean, rstd = K.mhla_gate_ln_fwd(vc, logits, g_, b_, eps)
        ctx.save_for_backward(vc, w1, w2, pre1, h, p, mean, rstd, gamma)
        ctx.biases = (b1, b2)
        ctx.vdt = v.dtype
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        vc, w1, w2, pre1, h, p, mean, rstd, gamma = ctx.saved_tensors
        Bn, Lr, E = vc.shape
        Hh = w2.shape[0]
        g_ = None if gamma is None else gamma.detach()
        dv_direct, dlogits, dg, db = K.mhla_gate_ln_bwd(K.to_compute(gy), vc, p, mean, rstd, g_)
        dl2 = dlogits.view(-1, Hh)
        b1, b2 = ctx.biases
        dw2, db2 = _wbgrad(w2, b2, dl2, h)
        w2c = shadow(w2)
        if K.smallk_mul_ok(dl2, w2c, pre1):          # 8 heads: a K = 8 contraction, HBM-bound (dl_smallk_mul)
            dpre1 = K.smallk_mul(dl2, w2c, pre1)
        else:
            dpre1 = K.mm(dl2, w2c, tb=True, mul_aux=pre1, mul_mode=K.MUL_VALUE)
        dw1, db1 = _wbgrad(w1, b1, dpre1, vc.view(-1, E))
        dv = K.mm(dpre1, shadow(w1), tb=True, res=dv_direct.view(-1, E)).view(Bn, Lr, E)
        return _back(dv, ctx.vdt), dw1, db1, dw2, db2, dg, db, None


# ================================================================================ GCN pieces
class SpmmFn(Function):
    """Degree-normalised neighbourhood sum of GraphConv (model/basic_model.py:596-630)."""

    @staticmethod
    def forward(ctx, h, graph):
        hc = K.to_compute(h)
        ctx.graph = graph
        ctx.hdt = h.dtype
        ctx.cdt = hc.dtype
        return K.spmm_norm(graph.indptr, graph.indices, graph.norm_src, graph.norm_dst, hc)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        g = ctx.graph
        dh = K.spmm_norm(g.indptr_t, g.indices_t, g.norm_dst, g.norm_src, _as(gy, ctx.cdt))
        return _back(dh, ctx.hdt), None


class ExpandVirtualFn(Function):
    """(R + 1, C) compact rows -> (N, C): real rows to their slots, the representative virtual row
    (the last one) to every other slot (graph.CompactMolGraph).  Backward: the real rows' gradients,
    and for the representative the SUM over all the slots it stands for."""

    @staticmethod
    def forward(ctx, xc, real_idx, n_full):
        R = real_idx.numel()
        out = xc[R].expand(n_full, xc.shape[1]).contiguous()
        out.index_copy_(0, real_idx, xc[:R])
        ctx.save_for_backward(real_idx)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (real_idx,) = ctx.saved_tensors
        g = g.contiguous()
        real = g.index_select(0, real_idx)
        virt = K.colsum(g) - K.colsum(real)                  # fp32 column sums (dl_colsum)
        return torch.cat((real, virt.to(g.dtype).unsqueeze(0))), None, None


class BatchNormFn(Function):
    """nn.BatchNorm1d over (rows, C) with the module's buffers updated in place."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, eps, momentum, training, relu_input=False,
                last_row_weight=1.0):
        """relu_input: x is the output of a ReLU whose own backward is skipped (Conv1dSameFn with
        mask_in_bwd=False); the mask (x > 0) is applied to dx inside the BatchNorm backward kernel.
        last_row_weight w > 1: the last row stands for w identical rows (dl_batchnorm_fwd)."""
        xc = K.to_compute(x)
        x2 = xc.view(-1, xc.shape[-1])
        g_ = None if gamma is None else gamma.detach()
        b_ = None if beta is None else beta.detach()
        y, mean, rstd = K.batchnorm_fwd(x2, g_, b_, running_mean, running_var, nbt, eps, momentum, training,
                                        last_row_weight if training else 1.0)
        ctx.save_for_backward(x2, gamma, beta, mean, rstd)
        ctx.meta = (training, x.dtype, x.shape, bool(relu_input), float(last_row_weight) if training else 1.0)
        return y.view(xc.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2, gamma, beta, mean, rstd = ctx.saved_tensors
        training, xdt, xshape, relu, lrw = ctx.meta
        g_ = None if gamma is None else gamma.detach()
        tg, tb = _grad_target(gamma), _grad_target(beta)
        if tg is not None and tb is not None:       # straight into the flat gradient buffer
            dx, _, _ = K.batchnorm_bwd(_as(gy, x2.dtype).view(x2.shape), x2, g_, mean, rstd, training,
                                       acc_into=(tg, tb), relu_mask=relu, last_row_weight=lrw)
            return _back(dx, xdt, xshape), None, None, None, None, None, None, None, None, None, None
        dx, dg, db = K.batchnorm_bwd(_as(gy, x2.dtype).view(x2.shape), x2, g_, mean, rstd, training,
                                     need_param_grads=gamma is not None, relu_mask=relu, last_row_weight=lrw)
        return _back(dx, xdt, xshape), dg, db, None, None, None, None, None, None, None, None


_bn_frozen = False


@contextlib.contextmanager
def frozen_bn_buffers():
    """Inside, train-mode BatchNorms normalise with batch statistics as usual but leave their
    running buffers alone (a feature-only pre-pass must not count a batch twice)."""
    global _bn_frozen
    prev, _bn_frozen = _bn_frozen, True
    try:
        yield
    finally:
        _bn_frozen = prev


def batch_norm(x, bn: torch.nn.BatchNorm1d, relu_input: bool = False, last_row_weight: float = 1.0):
    """Apply an nn.BatchNorm1d module's parameters/buffers with the dl_batchnorm kernels."""
    training = bn.training or bn.running_mean is None
    momentum = 0.1 if bn.momentum is None else bn.momentum
    update = training and not _bn_frozen
    return BatchNormFn.apply(x, bn.weight, bn.bias, bn.running_mean if (update or not training) else None,
                             bn.running_var if (update or not training) else None,
                             bn.num_batches_tracked if update else None, bn.eps, momentum, training,
                             relu_input, last_row_weight)


# ================================================================================ glue
class SitePoolFn(Function):
    @staticmethod
    def forward(ctx, x, S):
        xc = K.to_compute(x)
        ctx.meta = (S, x.dtype)
        return K.site_pool_fwd(xc, S)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        S, xdt = ctx.meta
        return _back(K.site_pool_bwd(K.to_compute(gy), S), xdt), None


def seq_mean(x: torch.Tensor) -> torch.Tensor:
    """torch.mean(x, dim=1) of (B, L, C) -> (B, C) with the dl_site_pool kernels.  A mean over a long
    sequence is taken in two stages (16 interleaved groups, then the 16 partial means): one stage would
    leave B*C/4 threads looping serially over L strided rows."""
    B, Lr, C_ = x.shape
    if Lr % 16 == 0 and Lr > 16:
        x = SitePoolFn.apply(x, Lr // 16)                 # (B, 16, C): mean over s of x[b, s*16 + j]
        Lr = 16
    return SitePoolFn.apply(x, Lr).view(B, C_)


class AddPEFn(Function):
    """dropout(x + pe) (model/PMMA/embed.py:51-52)."""

    @staticmethod
    def forward(ctx, x, pe, p, seed):
        xc = K.to_compute(x)
        ctx.meta = (p, seed, x.dtype, pe.shape)
        return K.add_pe(xc, pe.detach().contiguous(), p, seed)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        p, seed, xdt, pshape = ctx.meta
        g = K.act_bwd(K.to_compute(gy), None, K.ACT_NONE, (p, seed))
        n = 1
        for s in pshape:
            n *= s
        dpe = K.colsum(g.view(-1, n)).view(pshape)
        return _back(g, xdt), dpe, None, None


class EmbedFillFn(Function):
    """ProteinCNN's input: cat(embedding(tokens), fill_mask) in the compute dtype
    (model/basic_model.py:171-173) as one gather kernel; the backward sums the gradient rows per
    token in shared memory instead of nn.Embedding's sort-based dense backward."""

    @staticmethod
    def forward(ctx, tokens, fill, table, padding_idx):
        tok = tokens if tokens.dtype in (torch.int64, torch.float64) else tokens.long()
        tok = tok.contiguous()
        out = K.embed_fill_fwd(tok, fill.float().contiguous(), table.detach().contiguous())
        ctx.save_for_backward(tok, table)
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        tok, table = ctx.saved_tensors
        g = K.to_compute(gy)
        tgt = _grad_target(table)
        if tgt is not None:
            K.embed_fill_bwd(tok, g, tgt, ctx.padding_idx)
            return None, None, None, None
        dt = torch.zeros(table.shape, dtype=torch.float32, device=g.device)
        K.embed_fill_bwd(tok, g, dt, ctx.padding_idx)
        return None, None, dt, None


class Conv1dSameFn(Function):
    """nn.Conv1d(padding='same') + optional ReLU on channels-last activations, as an implicit GEMM
    on dl_gemm (TMA row-shifted A tiles, out-of-range rows zero-filled): the three convolutions of
    ProteinCNN (model/basic_model.py:163-178).  x (B, L, Cin) -> (B, L, Cout); w is the module's
    (Cout, Cin, k) parameter.  PyTorch's 'same' puts the extra pad of an even kernel on the right."""

    @staticmethod
    def forward(ctx, x, w, b, relu, mask_in_bwd=True):
        """mask_in_bwd=False: the consumer (batch_norm(..., relu_input=True)) applies the ReLU mask
        to the gradient it sends back, so this backward takes the gradient as is."""
        xc = K.to_compute(x)
        Cout, Cin, k = w.shape
        left = (k - 1) // 2
        wc = shadow(w)
        w_taps = wc.permute(0, 2, 1).reshape(Cout, k * Cin).contiguous()
        out = torch.empty(xc.shape[:2] + (Cout,), dtype=xc.dtype, device=xc.device)
        K.conv1d_same(xc, w_taps, out, taps=k, left=left, bias=None if b is None else b.detach(),
                      act=K.ACT_RELU if relu else K.ACT_NONE)
        # the tap-reversed, transposed weight of the dX convolution only depends on the weights: it is laid
        # out now, on a side stream, instead of in the middle of the backward's longest chain
        wd, wd_ready = None, None
        if _needs_backward(ctx) and ctx.needs_input_grad[0]:
            side = branch_stream(xc)
            cur = torch.cuda.current_stream()
            if side is not None and side != cur:
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    wd = wc.flip(2).permute(1, 2, 0).reshape(Cin, k * Cout).contiguous()
                    wd_ready = side.record_event()
            else:
                wd = wc.flip(2).permute(1, 2, 0).reshape(Cin, k * Cout).contiguous()
        ctx.save_for_backward(xc, w, out if (relu and mask_in_bwd) else None, b, wd)
        ctx.wd_ready = wd_ready
        ctx.meta = (relu and mask_in_bwd, left, x.dtype, b is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xc, w, out, b, wd = ctx.saved_tensors
        relu, left, xdt, has_b = ctx.meta
        Cout, Cin, k = w.shape
        g = K.to_compute(gy)
        if relu:
            g = K.act_bwd(g, out, K.ACT_RELU)          # mask from the post-ReLU output (y > 0 <=> pre > 0)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if wd is None:
                wd = shadow(w).flip(2).permute(1, 2, 0).reshape(Cin, k * Cout).contiguous()
            elif ctx.wd_ready is not None:
                torch.cuda.current_stream().wait_event(ctx.wd_ready)
                crosses(torch.cuda.current_stream(), wd)
            dx = torch.empty_like(xc)
            K.conv1d_same(g, wd, dx, taps=k, left=k - 1 - left)
            dx = _back(dx, xdt)
        if ctx.needs_input_grad[1]:
            tgt = _grad_target(w)
            if tgt is not None:
                # (taps, Cout, Cin) partial on the deferred stream, then added into the (Cout, Cin, k) gradient
                with _param_grad_stream(g, xc):
                    tgt.add_(K.conv1d_same_wgrad(g, xc, k, left).permute(1, 2, 0))
            else:
                dw = K.conv1d_same_wgrad(g, xc, k, left).permute(1, 2, 0).contiguous()
        if has_b and ctx.needs_input_grad[2]:
            db = _bgrad(b, g.view(-1, Cout))
        return dx, dw, db, None, None


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    """(..., C) -> a (rows, C) view with unit inner stride and ONE row stride (a copy only when the leading
    dimensions do not collapse)."""
    C_ = t.shape[-1]
    if t.stride(-1) == 1:
        try:
            return t.view(-1, C_)
        except RuntimeError:
            pass
    return t.contiguous().view(-1, C_)


class CatLastFn(Function):
    """torch.cat((a, b), dim=-1) (model/DrugLAMP.py:57,66; model/PMMA/encoder.py:46) with dl_copy_rows: two
    16-byte-vector row copies forward, and two CONTIGUOUS halves of the gradient backward (autograd's own
    backward hands out column-slice views, which every consumer then compacts with an element-wise copy)."""

    @staticmethod
    def forward(ctx, a, b):
        Ca, Cb = a.shape[-1], b.shape[-1]
        a2, b2 = _rows2d(a), _rows2d(b)
        out = torch.empty((a2.shape[0], Ca + Cb), dtype=a.dtype, device=a.device)
        K.copy_rows(a2, out[:, :Ca])
        K.copy_rows(b2, out[:, Ca:])
        ctx.meta = (a.shape, b.shape)
        return out.view(*a.shape[:-1], Ca + Cb)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        ashape, bshape = ctx.meta
        Ca, Cb = ashape[-1], bshape[-1]
        g2 = _rows2d(g)
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = K.copy_rows(g2[:, :Ca], torch.empty((g2.shape[0], Ca), dtype=g.dtype, device=g.device)).view(ashape)
        if ctx.needs_input_grad[1]:
            gb = K.copy_rows(g2[:, Ca:], torch.empty((g2.shape[0], Cb), dtype=g.dtype, device=g.device)).view(bshape)
        return ga, gb


def cat_last(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """cat((a, b), -1); falls back to torch.cat when the shapes do not fit the 16-byte row copies."""
    es = a.element_size()
    ok = (a.is_cuda and a.dtype == b.dtype and a.shape[:-1] == b.shape[:-1] and (a.shape[-1] * es) % 16 == 0
          and (b.shape[-1] * es) % 16 == 0)
    return CatLastFn.apply(a, b) if ok else torch.cat((a, b), dim=-1)


class TransposeFn(Function):
    """(B, R, C) -> (B, C, R) contiguous."""

    @staticmethod
    def forward(ctx, x):
        ctx.xdt = x.dtype
        return K.transpose_last2(K.to_compute(x))

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        return _back(K.transpose_last2(K.to_compute(gy)), ctx.xdt)


class CnnTailFn(Function):
    """The tail of ProteinCNN on the channels-last activation x (B, L, C) = relu(conv3(.)):
    BatchNorm1d over the channels, the (B, C, L) layout of the reference and its ``.view(B, L, C)``
    reinterpretation (model/basic_model.py:178-179) -- statistics, then ONE
    normalise-and-transpose pass (dl_bn_transpose).  With site_len S > 0 the site mean of
    model/DrugLAMP.py:35-37 follows inside the same Function, so the backward maps the pooled gradient
    straight to the channels-last layout (dl_site_pool_view_bwd) instead of expanding it 9x and transposing it.
    relu_input: x is a ReLU output whose own backward is skipped; its mask rides on the BatchNorm backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, eps, momentum, training, relu_input, S):
        xc = K.to_compute(x)
        B, Lr, C_ = xc.shape
        x2 = xc.view(-1, C_)
        g_ = None if gamma is None else gamma.detach()
        b_ = None if beta is None else beta.detach()
        mean, rstd = K.batchnorm_stats(x2, running_mean, running_var, nbt, eps, momentum, training)
        y = K.bn_transpose(xc, mean, rstd, g_, b_).view(B, Lr, C_)          # the reinterpreting view
        ctx.save_for_backward(x2, gamma, beta, mean, rstd)
        ctx.meta = (training, x.dtype, x.shape, bool(relu_input), int(S))
        return K.site_pool_fwd(y, S) if S > 0 else y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2, gamma, beta, mean, rstd = ctx.saved_tensors
        training, xdt, xshape, relu, S = ctx.meta
        B, Lr, C_ = xshape
        g = _as(gy, x2.dtype)
        if S > 0:
            dy = K.site_pool_view_bwd(g, S)                                   # (B, L/S, C) -> (B, L, C)
        else:
            dy = K.bn_transpose(g.view(B, C_, Lr))                            # (B, C, L) -> (B, L, C)
        g_ = None if gamma is None else gamma.detach()
        tg, tb = _grad_target(gamma), _grad_target(beta)
        nones = (None,) * 8
        if tg is not None and tb is not None:       # straight into the flat gradient buffer
            dx, _, _ = K.batchnorm_bwd(dy.view(x2.shape), x2, g_, mean, rstd, training, acc_into=(tg, tb),
                                       relu_mask=relu)
            return (_back(dx, xdt, xshape), None, None) + nones
        dx, dg, db = K.batchnorm_bwd(dy.view(x2.shape), x2, g_, mean, rstd, training,
                                     need_param_grads=gamma is not None, relu_mask=relu)
        return (_back(dx, xdt, xshape), dg, db) + nones


def cnn_tail_ok(x: torch.Tensor, S: int) -> bool:
    """Shapes dl_bn_transpose / dl_site_pool_view_bwd serve (16-byte vectors along both tile edges)."""
    v = 8 if K.compute_dtype() == torch.bfloat16 else 4
    return x.is_cuda and x.dim() == 3 and x.shape[1] % v == 0 and x.shape[2] % v == 0 and (S <= 0 or x.shape[1] % S == 0)


def cnn_tail(x, bn: torch.nn.BatchNorm1d, relu_input: bool = True, site_len: int = 0):
    training = bn.training or bn.running_mean is None
    momentum = 0.1 if bn.momentum is None else bn.momentum
    update = training and not _bn_frozen
    keep = update or not training
    return CnnTailFn.apply(x, bn.weight, bn.bias, bn.running_mean if keep else None, bn.running_var if keep else None,
                           bn.num_batches_tracked if update else None, bn.eps, momentum, training, relu_input,
                           site_len)


class ActFn(Function):
    @staticmethod
    def forward(ctx, x, act):
        xc = K.to_compute(x)
        ctx.save_for_backward(xc)
        ctx.meta = (act, x.dtype)
        return K.act_fwd(xc, act)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (xc,) = ctx.saved_tensors
        act, xdt = ctx.meta
        return _back(K.act_bwd(K.to_compute(gy), xc, act), xdt), None


class L2NormFn(Function):
    """l2norm = F.normalize(t, dim=-1) (utils.py:443-444)."""

    @staticmethod
    def forward(ctx, x):
        xc = K.to_compute(x)
        y, norm = K.l2norm_fwd(xc)
        ctx.save_for_backward(y, norm)
        ctx.xdt = x.dtype
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        y, norm = ctx.saved_tensors
        return _back(K.l2norm_bwd(K.to_compute(gy), y, norm), ctx.xdt)


class CMTripletFn(Function):
    @staticmethod
    def forward(ctx, cos, G, margin):
        c32 = K.cast(cos.contiguous(), torch.float32)
        loss, acc = K.cm_triplet_fwd(c32, G, margin)
        ctx.save_for_backward(c32, G, acc)
        ctx.meta = (margin, cos.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gl):
        c32, G, acc = ctx.saved_tensors
        margin, cdt = ctx.meta
        dcos = K.cm_triplet_bwd(c32, G, margin, acc, gl.contiguous().float())
        return _back(dcos, cdt), None, None


class BCEFn(Function):
    @staticmethod
    def forward(ctx, score, y):
        s32 = K.cast(score.contiguous(), torch.float32).view(-1)
        y32 = y.to(torch.float32).contiguous()
        prob, loss = K.bce_fwd(s32, y32)
        ctx.save_for_backward(prob, y32)
        ctx.meta = (score.dtype, score.shape)
        ctx.mark_non_differentiable(prob)
        return prob, loss

    @staticmethod
    @once_differentiable
    def backward(ctx, _gp, gl):
        prob, y32 = ctx.saved_tensors
        sdt, sshape = ctx.meta
        ds = K.bce_bwd(prob, y32, gl.contiguous().float())
        return _back(ds, sdt, sshape), None


class CrossEntropyFn(Function):
    """F.cross_entropy(logits (..., C), labels (...), ignore_index) with mean reduction."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        C = logits.shape[-1]
        x = K.to_compute(logits).view(-1, C)
        lab = labels.reshape(-1).to(torch.int64).contiguous()
        loss, acc = K.cross_entropy_fwd(x, lab, C, ignore_index)
        ctx.save_for_backward(x, lab, acc)
        ctx.meta = (C, ignore_index, logits.dtype, logits.shape)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gl):
        x, lab, acc = ctx.saved_tensors
        C, ignore_index, ldt, lshape = ctx.meta
        dx, _ = K.cross_entropy_bwd(x, lab, C, acc, gl.contiguous().float(), ignore_index)
        return _back(dx, ldt, lshape), None, None


class MatmulNTFn(Function):
    """C = A B^T for row-major A (M,K), B (N,K) -- the CrossModality similarity matrix."""

    @staticmethod
    def forward(ctx, a, b):
        M, N = a.shape[0], b.shape[0]
        al = _align()
        Mp, Np = -(-M // al) * al, -(-N // al) * al
        pad = torch.nn.functional.pad
        ac, bc = K.to_compute(a), K.to_compute(b)
        if Mp != M:
            ac = pad(ac, (0, 0, 0, Mp - M))
        if Np != N:
            bc = pad(bc, (0, 0, 0, Np - N))
        ctx.save_for_backward(ac, bc)
        ctx.meta = (a.dtype, b.dtype, M, N)
        return K.mm(ac, bc, out_dtype=torch.float32)[:M, :N].contiguous()

    @staticmethod
    @once_differentiable
    def backward(ctx, gc):
        ac, bc = ctx.saved_tensors
        adt, bdt, M, N = ctx.meta
        Mp, Np = ac.shape[0], bc.shape[0]
        g = K.cast(torch.nn.functional.pad(gc, (0, Np - N, 0, Mp - M)).contiguous(), ac.dtype)
        da = K.mm(g, bc, tb=True)[:M]           # (M,N) @ (N,K)
        db = K.mm(g, ac, ta=True, tb=True)[:N]  # (N,M) @ (M,K)
        return _back(da, adt), _back(db, bdt)
