"""Thin, autograd-free Python wrappers over the C ABI.

Every function takes CUDA tensors (2-D operands may be row/column *views* of larger buffers: only
``stride(-1) == 1`` is required, the row stride becomes the leading dimension) and launches
hand-written sm_100a kernels from ``libdruglamp_sm100.so`` on the current stream.  Nothing here
falls back to PyTorch arithmetic.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, MUL_GELU_GRAD, MUL_NONE, MUL_RELU_MASK, MUL_VALUE  # noqa: F401

_compute_dtype = torch.float32


def set_compute_dtype(dtype: torch.dtype) -> None:
    """fp32 (TF32 tensor cores, the reference's own matmul precision, main.py:43) or bf16."""
    global _compute_dtype
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("compute dtype must be torch.float32 or torch.bfloat16")
    _compute_dtype = dtype


def compute_dtype() -> torch.dtype:
    return _compute_dtype


class local_compute_dtype:
    """``with local_compute_dtype(torch.float32): ...`` -- run a sub-graph's FORWARD in another
    compute dtype (its Functions keep the dtype of what they saved for their backward).  When the
    surrounding mode is bf16, the fp32 island's GEMMs run as plain TF32 (one MMA pass, fp32
    accumulation and fp32 storage -- the precision the reference itself selects, main.py:43) instead of
    the 3xTF32 split of the fp32 parity mode; LinearFn remembers the choice for its backward."""

    def __init__(self, dtype, precise=None):
        self.dtype, self.precise = dtype, precise

    def __enter__(self):
        global _compute_dtype
        self.prev, self.prev_precise = _compute_dtype, L.FP32_PRECISE
        if self.dtype is not None:
            if self.dtype == torch.float32 and _compute_dtype == torch.bfloat16:
                L.FP32_PRECISE = False if self.precise is None else bool(self.precise)
            _compute_dtype = self.dtype

    def __exit__(self, *exc):
        global _compute_dtype
        _compute_dtype = self.prev
        L.FP32_PRECISE = self.prev_precise
        return False


class fp32_precise:
    """Temporarily select 3xTF32 (True) or plain TF32 (False) for fp32 GEMM operands."""

    def __init__(self, flag):
        self.flag = flag

    def __enter__(self):
        self.prev = L.FP32_PRECISE
        L.FP32_PRECISE = self.flag

    def __exit__(self, *exc):
        L.FP32_PRECISE = self.prev
        return False


def set_dropout_step(counter: Optional[torch.Tensor]) -> None:
    """A device-resident int64 step counter (or None) that every dropout launch mixes into its seed
    ON THE DEVICE: a captured CUDA graph then draws a fresh mask on every replay, while the forward
    and backward of one step (the counter moves in the optimiser update) regenerate the same one."""
    if counter is not None and (counter.dtype != torch.int64 or counter.numel() != 1 or not counter.is_cuda):
        raise TypeError("the dropout step counter must be a CUDA int64 scalar tensor")
    L.DROPOUT_STEP = counter


def _ld(t: torch.Tensor) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def cast(x: torch.Tensor, dtype: torch.dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dtype conversion of a contiguous tensor with the dl_cast kernel."""
    if x.dtype == dtype and out is None:
        return x
    if not x.is_contiguous():
        x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=dtype, device=x.device)
    L.call("dl_cast", x.data_ptr(), L.dt(x), out.data_ptr(), L.dt(out), x.numel())
    return out


def to_compute(x: torch.Tensor) -> torch.Tensor:
    """Contiguous tensor in the compute dtype."""
    if not x.is_contiguous():
        x = x.contiguous()
    return cast(x, _compute_dtype)


# ----------------------------------------------------------------------------- GEMM flavours
def mm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, ta: bool = False,
       tb: bool = False, bias=None, act: int = ACT_NONE, pre: Optional[torch.Tensor] = None,
       mul_aux=None, mul_mode: int = MUL_NONE, res: Optional[torch.Tensor] = None,
       drop: Tuple[float, int] = (0.0, 0), alpha: float = 1.0, out_dtype=None,
       tile_n: int = 0, accumulate: bool = False, colsum_a: Optional[torch.Tensor] = None,
       pre_mode: int = 0) -> torch.Tensor:
    """out[M,N] = epilogue(alpha * op(a) @ op(b)).  colsum_a (fp32 [M], bf16 operands, ta=True):
    += column sums of a, i.e. the bias gradient when this is a weight-gradient GEMM.

    ta=False: a is [M,K];  ta=True: a is stored [K,M].
    tb=False: b is [N,K] (nn.Linear weight layout);  tb=True: b is stored [K,N].
    """
    M, K = (a.shape[1], a.shape[0]) if ta else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if tb else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"contraction mismatch: {tuple(a.shape)} ta={ta} vs {tuple(b.shape)} tb={tb}")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or a.dtype, device=a.device)
    ldr = 0
    if res is not None and res.data_ptr() != out.data_ptr():
        ldr = _ld(res)
    L.gemm(a, b, out, M=M, N=N, K=K, lda=_ld(a), ldb=_ld(b), ldc=_ld(out), trans_a=ta, trans_b=tb,
           alpha=alpha, bias=bias, act=act, preact_out=pre, mul_aux=mul_aux, mul_mode=mul_mode,
           residual=res, ldr=ldr, drop_p=drop[0], drop_seed=drop[1], tile_n=tile_n, accumulate=accumulate,
           colsum_a=colsum_a, pre_mode=pre_mode)
    return out


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    if out is None:
        out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    L.call("dl_colsum", x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _ld(x), int(accumulate), L.dt(x))
    return out


def act_bwd(dy: torch.Tensor, pre: Optional[torch.Tensor], act: int, drop=(0.0, 0)) -> torch.Tensor:
    """dy * act'(pre) * dropout_mask on contiguous tensors."""
    if act == ACT_NONE and drop[0] == 0.0:
        return dy
    g = torch.empty_like(dy)
    L.call("dl_act_bwd", dy.data_ptr(), None if pre is None else pre.data_ptr(), g.data_ptr(),
           dy.numel(), act, drop[0], drop[1], L.ptr(L.DROPOUT_STEP), L.dt(dy))
    return g


def act_fwd(x: torch.Tensor, act: int) -> torch.Tensor:
    y = torch.empty_like(x)
    L.call("dl_act_fwd", x.data_ptr(), y.data_ptr(), x.numel(), act, L.dt(x))
    return y


def l2norm_fwd(x: torch.Tensor, eps: float = 1e-12):
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty_like(x)
    norm = torch.empty(rows, dtype=torch.float32, device=x.device)
    L.call("dl_l2norm_fwd", x.data_ptr(), y.data_ptr(), norm.data_ptr(), rows, cols, eps, L.dt(x))
    return y, norm


def l2norm_bwd(dy, y, norm):
    rows, cols = y.numel() // y.shape[-1], y.shape[-1]
    dx = torch.empty_like(y)
    L.call("dl_l2norm_bwd", dy.data_ptr(), y.data_ptr(), norm.data_ptr(), dx.data_ptr(), rows, cols, L.dt(y))
    return dx


def dropout(x: torch.Tensor, p: float, seed: int) -> torch.Tensor:
    if p == 0.0:
        return x
    y = torch.empty_like(x)
    L.call("dl_dropout", x.data_ptr(), y.data_ptr(), x.numel(), p, seed, L.ptr(L.DROPOUT_STEP), L.dt(x))
    return y


def add_pe(x: torch.Tensor, pe: torch.Tensor, p: float = 0.0, seed: int = 0) -> torch.Tensor:
    y = torch.empty_like(x)
    L.call("dl_add_pe", x.data_ptr(), pe.data_ptr(), y.data_ptr(), x.numel(), pe.numel(), p, seed,
           L.ptr(L.DROPOUT_STEP), L.dt(x))
    return y


# ----------------------------------------------------------------------------- row kernels
def layernorm_fwd(x: torch.Tensor, gamma, beta, eps: float, save: bool = True):
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty_like(x)
    mean = rstd = None
    if save:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
    L.call("dl_layernorm_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
           L.ptr(mean), L.ptr(rstd), rows, cols, eps, L.dt(x))
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, need_param_grads: bool = True, acc_into=None, dx_add=None):
    """acc_into = (dgamma, dbeta) fp32 gradient buffers to ADD the parameter gradients to.
    dx_add: a tensor shaped like x that is added to dx (the skip connection's gradient)."""
    rows, cols = x.numel() // x.shape[-1], x.shape[-1]
    dx = torch.empty_like(x)
    dg = db = None
    if acc_into is not None:
        dg, db = acc_into
    elif need_param_grads:
        dg = torch.empty(cols, dtype=torch.float32, device=x.device)
        db = torch.empty_like(dg)
    L.call("dl_layernorm_bwd", dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(),
           rstd.data_ptr(), dx.data_ptr(), L.ptr(dx_add), L.ptr(dg), L.ptr(db), rows, cols,
           int(acc_into is not None), L.dt(x))
    return dx, dg, db


def _row_ld(t: torch.Tensor) -> int:
    """Row stride of a (..., rows, cols) tensor whose leading dims collapse onto evenly spaced rows."""
    if t.dim() < 2:
        return t.shape[-1]
    ld = t.stride(-2)
    assert t.stride(-1) == 1 and ld >= t.shape[-1]
    run = ld
    for size, stride in zip(reversed(t.shape[:-1]), reversed(t.stride()[:-1])):
        assert size == 1 or stride == run, "softmax rows must be evenly spaced"
        run *= size
    return ld


def softmax_fwd(s: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax over the last dim (in place when out is None).  The rows may be padded: the row
    stride is ``s.stride(-2)`` and ``out`` must share the layout."""
    cols = s.shape[-1]
    out = s if out is None else out
    ld = _row_ld(s)
    assert out.stride() == s.stride()
    L.call("dl_softmax_fwd", s.data_ptr(), out.data_ptr(), s.numel() // cols, cols, ld, L.dt(s))
    return out


def softmax_bwd(p: torch.Tensor, dp: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """in place on dp: dp <- scale * p * (dp - sum(p * dp))."""
    cols = p.shape[-1]
    assert dp.stride() == p.stride()
    L.call("dl_softmax_bwd", p.data_ptr(), dp.data_ptr(), dp.data_ptr(), p.numel() // cols, cols,
           _row_ld(p), scale, L.dt(p))
    return dp


# ----------------------------------------------------------------------------- fused attention
ATTN_MAX_KEYS = 512


def attn_supported(q: torch.Tensor, k: torch.Tensor, H: int) -> bool:
    """dl_attn_* serves bf16, head dim 64 / 128 and up to 512 keys (what DrugLAMP uses); other
    shapes run the same math as separate dl_gemm / dl_softmax launches."""
    d = q.shape[-1] // H
    return q.dtype == torch.bfloat16 and d in (64, 128) and k.shape[1] <= ATTN_MAX_KEYS


def _attn_args(q, k, v, o, lse, H, scale):
    S2, B, Lq, HD = q.shape
    a = L.AttnArgs()
    a.q, a.k, a.v, a.o, a.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr()
    a.B, a.H, a.S2, a.Lq, a.Lk, a.d = B, H, S2, Lq, k.shape[1], HD // H
    a.q_ld, a.q_sb, a.q_ss = q.stride(2), q.stride(1), q.stride(0) if S2 > 1 else 0
    a.k_ld, a.k_sb, a.v_ld, a.v_sb = k.stride(1), k.stride(0), v.stride(1), v.stride(0)
    a.o_ld, a.o_sb, a.o_ss = o.stride(1), o.stride(0), HD
    a.scale = scale
    for t in (q, k, v, o):
        if t.stride(-1) != 1 or t.dtype != torch.bfloat16:
            raise ValueError("dl_attn needs bf16 operands with unit inner stride")
    return a


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, H: int, scale: float, want_raw: bool = False,
             out: Optional[torch.Tensor] = None):
    """q (S2, B, Lq, H*d), k / v (B, Lk, H*d) bf16 views (unit inner stride, the rest free).
    -> O (B, Lq, S2*H*d), lse (S2, B, H, Lq) fp32 [log2 domain], raw (B, H, Lq, Lk) | None.
    `out`: a preallocated (B, Lq, S2*H*d) destination with free row / pair strides (e.g. a view of a
    sequence-first buffer)."""
    S2, B, Lq, HD = q.shape
    Lk = k.shape[1]
    o = out if out is not None else torch.empty((B, Lq, S2 * HD), dtype=q.dtype, device=q.device)
    if tuple(o.shape) != (B, Lq, S2 * HD):
        raise ValueError("attn_fwd: out must be (B, Lq, S2*H*d)")
    lse = torch.empty((S2, B, H, Lq), dtype=torch.float32, device=q.device)
    a = _attn_args(q, k, v, o, lse, H, scale)
    raw = None
    if want_raw:
        raw = torch.empty((B, H, Lq, Lk), dtype=q.dtype, device=q.device)
        a.raw, a.raw_ld = raw.data_ptr(), Lk
    L.check(L.lib().dl_attn_fwd(L.C.byref(a), L.stream_ptr()), "dl_attn_fwd")
    return o, lse, raw


def attn_bwd(d_o: torch.Tensor, q, k, v, o, lse, H: int, scale: float, dq: torch.Tensor, dk: torch.Tensor,
             dv: torch.Tensor, dq_accumulate: bool = False) -> None:
    """Gradients of attn_fwd into the (strided) destinations dq (like q), dk, dv (like k, v)."""
    S2, B, Lq, HD = q.shape
    if d_o.stride() != o.stride() or d_o.dtype != o.dtype:
        raise ValueError("dl_attn_bwd: d_o must share o's layout")
    a = _attn_args(q, k, v, o, lse, H, scale)
    a.d_o = d_o.data_ptr()
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.dq_ld, a.dq_sb, a.dq_ss = dq.stride(2), dq.stride(1), dq.stride(0) if S2 > 1 else 0
    a.dk_ld, a.dk_sb, a.dv_ld, a.dv_sb = dk.stride(1), dk.stride(0), dv.stride(1), dv.stride(0)
    dvec = torch.empty((S2, B, H, Lq), dtype=torch.float32, device=q.device)
    a.dvec = dvec.data_ptr()
    a.dq_accumulate = int(dq_accumulate)
    L.check(L.lib().dl_attn_bwd(L.C.byref(a), L.stream_ptr()), "dl_attn_bwd")


def smallk_mul_ok(g: torch.Tensor, w: torch.Tensor, aux: torch.Tensor) -> bool:
    K_, N = w.shape
    return (g.is_cuda and g.dtype == w.dtype == aux.dtype == torch.bfloat16 and K_ <= 16 and N % 8 == 0
            and g.dim() == 2 and g.shape[1] == K_ and g.stride(1) == 1 and g.stride(0) % 8 == 0
            and g.stride(0) >= (8 if K_ <= 8 else 16) and w.is_contiguous() and aux.is_contiguous()
            and all(t.data_ptr() % 16 == 0 for t in (g, w, aux)))


def smallk_mul(g: torch.Tensor, w: torch.Tensor, aux: torch.Tensor) -> torch.Tensor:
    """(g [M, K] @ w [K, N]) * aux [M, N] for K <= 16 (dl_smallk_mul): HBM-bound, no tensor cores."""
    out = torch.empty_like(aux)
    L.call("dl_smallk_mul", g.data_ptr(), w.data_ptr(), aux.data_ptr(), out.data_ptr(), g.shape[0], w.shape[1],
           w.shape[0], g.stride(0))
    return out


# ----------------------------------------------------------------------------- fused FFN
FFN_WIDTH = 256      # model width the fused feed-forward kernels are built for


def ffn_supported(x2: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor) -> bool:
    """dl_ffn_fwd / dl_ffn_bwd serve bf16 rows of width 256 with a hidden width that is a multiple of 128."""
    return (x2.is_cuda and x2.dtype == torch.bfloat16 and x2.dim() == 2 and x2.shape[1] == FFN_WIDTH
            and x2.stride(1) == 1 and x2.stride(0) % 16 == 0 and x2.data_ptr() % 32 == 0
            and tuple(w1.shape) == (w2.shape[1], FFN_WIDTH) and w2.shape[0] == FFN_WIDTH and w1.shape[0] % 128 == 0
            and w1.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16 and w1.is_contiguous() and w2.is_contiguous())


def _ffn_args(x2, w1, w2, hidden, dact, y) -> "L.FfnArgs":
    a = L.FfnArgs()
    a.x, a.w1, a.w2, a.y = x2.data_ptr(), w1.data_ptr(), w2.data_ptr(), y.data_ptr()
    a.hidden = None if hidden is None else hidden.data_ptr()
    a.dact = None if dact is None else dact.data_ptr()
    a.M, a.D, a.Dh = x2.shape[0], x2.shape[1], w1.shape[0]
    a.ldx, a.ldy = x2.stride(0), y.stride(0)
    a.ldh = w1.shape[0] if hidden is None else hidden.stride(0)
    for t in (hidden, dact):
        if t is not None and (t.dtype != torch.bfloat16 or t.stride(1) != 1 or t.stride(0) != a.ldh):
            raise ValueError("dl_ffn: hidden / dact must be bf16 [M, Dh] with one row stride")
    return a


def ffn_fwd(x2: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
            res: Optional[torch.Tensor], drop: Tuple[float, int, int] = (0.0, 0, 0), keep: bool = True):
    """y = dropout(dropout(gelu(x2 w1^T + b1)) w2^T + b2) + res in ONE launch (dl_ffn_fwd).
    -> (y, hidden, dact); hidden / dact (the backward's operands) are None unless `keep`."""
    M, Dh = x2.shape[0], w1.shape[0]
    y = torch.empty((M, FFN_WIDTH), dtype=x2.dtype, device=x2.device)
    hidden = torch.empty((M, Dh), dtype=x2.dtype, device=x2.device) if keep else None
    dact = torch.empty_like(hidden) if keep else None
    a = _ffn_args(x2, w1, w2, hidden, dact, y)
    a.b1, a.b2 = b1.data_ptr(), b2.data_ptr()
    if res is not None:
        a.residual, a.ldr = res.data_ptr(), res.stride(0)
    p, s1, s2 = drop
    a.drop_p, a.seed1, a.seed2 = p, s1, s2
    a.drop_seed_step = L.ptr(L.DROPOUT_STEP) if p > 0 else None
    L.check(L.lib().dl_ffn_fwd(L.C.byref(a), L.stream_ptr()), "dl_ffn_fwd")
    if L.PROFILE is not None:
        L.PROFILE.append({"flops": 4.0 * M * Dh * FFN_WIDTH, "args": a, "fn": "dl_ffn_fwd",
                          "keep": (x2, w1, b1, w2, b2, res, y, hidden, dact), "shape": ("ffn_fwd", M, Dh)})
    return y, hidden, dact


def ffn_bwd(g2: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor, dact: torch.Tensor):
    """-> (dpre = (g2 w2) * dact  [M, Dh],  dx = dpre w1  [M, 256]) in ONE launch (dl_ffn_bwd)."""
    M = g2.shape[0]
    dpre = torch.empty_like(dact)
    dx = torch.empty((M, FFN_WIDTH), dtype=g2.dtype, device=g2.device)
    a = _ffn_args(g2, w1, w2, dpre, dact, dx)
    L.check(L.lib().dl_ffn_bwd(L.C.byref(a), L.stream_ptr()), "dl_ffn_bwd")
    if L.PROFILE is not None:
        L.PROFILE.append({"flops": 4.0 * M * w1.shape[0] * FFN_WIDTH, "args": a, "fn": "dl_ffn_bwd",
                          "keep": (g2, w1, w2, dact, dpre, dx), "shape": ("ffn_bwd", M, w1.shape[0])})
    return dpre, dx


# ----------------------------------------------------------------------------- GCN
def spmm_norm(indptr, indices, norm_src, norm_dst, h: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(h)
    L.call("dl_spmm_norm", indptr.data_ptr(), indices.data_ptr(), norm_src.data_ptr(),
           norm_dst.data_ptr(), h.data_ptr(), out.data_ptr(), h.shape[0], h.shape[1], L.dt(h))
    return out


def batchnorm_fwd(x: torch.Tensor, gamma, beta, running_mean, running_var, nbt, eps: float,
                  momentum: float, training: bool, last_row_weight: float = 1.0):
    rows, cols = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(cols, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    ws = torch.empty(2 * cols, dtype=torch.float64, device=x.device)
    L.call("dl_batchnorm_fwd", x.data_ptr(), L.ptr(gamma), L.ptr(beta), y.data_ptr(), mean.data_ptr(),
           rstd.data_ptr(), L.ptr(running_mean), L.ptr(running_var), L.ptr(nbt), ws.data_ptr(), rows, cols,
           eps, momentum, int(training), float(last_row_weight), L.dt(x))
    return y, mean, rstd


def batchnorm_stats(x: torch.Tensor, running_mean, running_var, nbt, eps: float, momentum: float, training: bool):
    """(mean, rstd) of nn.BatchNorm1d over (rows, cols) -- batch statistics in training (running buffers
    updated), running statistics otherwise -- without normalising (dl_batchnorm_fwd, y = NULL)."""
    rows, cols = x.shape
    mean = torch.empty(cols, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    ws = torch.empty(2 * cols, dtype=torch.float64, device=x.device)
    L.call("dl_batchnorm_fwd", x.data_ptr(), None, None, None, mean.data_ptr(), rstd.data_ptr(),
           L.ptr(running_mean), L.ptr(running_var), L.ptr(nbt), ws.data_ptr(), rows, cols, eps, momentum,
           int(training), 1.0, L.dt(x))
    return mean, rstd


def bn_transpose_ok(x: torch.Tensor) -> bool:
    v = 16 // x.element_size()
    return x.dim() == 3 and x.is_contiguous() and x.shape[1] % v == 0 and x.shape[2] % v == 0 and x.data_ptr() % 16 == 0


def bn_transpose(x: torch.Tensor, mean=None, rstd=None, gamma=None, beta=None) -> torch.Tensor:
    """(B, R, C) -> (B, C, R) contiguous, normalised per column c on the way when mean / rstd are given."""
    B, R, Cc = x.shape
    y = torch.empty((B, Cc, R), dtype=x.dtype, device=x.device)
    L.call("dl_bn_transpose", x.data_ptr(), y.data_ptr(), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(beta),
           B, R, Cc, L.dt(x))
    return y


def site_pool_view_bwd(g: torch.Tensor, S: int) -> torch.Tensor:
    """g (B, L/S, C) -> dx (B, L, C): gradient through transpose -> .view(B, L, C) -> site mean (dl_site_pool_view_bwd)."""
    B, P, C = g.shape
    g = g.contiguous()
    dx = torch.empty((B, P * S, C), dtype=g.dtype, device=g.device)
    L.call("dl_site_pool_view_bwd", g.data_ptr(), dx.data_ptr(), B, S, P * S, C, L.dt(g))
    return dx


def batchnorm_bwd(dy, x, gamma, mean, rstd, training: bool, need_param_grads: bool = True, acc_into=None,
                  relu_mask: bool = False, last_row_weight: float = 1.0):
    """acc_into = (dgamma, dbeta) fp32 gradient buffers to ADD the parameter gradients to.
    relu_mask: x is a ReLU output; dx is additionally multiplied by (x > 0)."""
    rows, cols = x.shape
    dx = torch.empty_like(x)
    ws = torch.empty(2 * cols, dtype=torch.float64, device=x.device)
    dg = db = None
    if acc_into is not None:
        dg, db = acc_into
    elif need_param_grads:
        dg = torch.empty(cols, dtype=torch.float32, device=x.device)
        db = torch.empty_like(dg)
    L.call("dl_batchnorm_bwd", dy.data_ptr(), x.data_ptr(), L.ptr(gamma), mean.data_ptr(), rstd.data_ptr(),
           dx.data_ptr(), L.ptr(dg), L.ptr(db), ws.data_ptr(), rows, cols, int(training),
           int(acc_into is not None), int(relu_mask), float(last_row_weight), L.dt(x))
    return dx, dg, db


# ----------------------------------------------------------------------------- glue
def fillbit_pool(x: torch.Tensor, S: int, want_bit=True, want_cat=False, want_pooled=True,
                 pooled_dtype=None, pad_to: int = 1):
    """x (B, S*L, C) fp32 -> (bit (B,S*L) | None, cat (B,S*L,C+1) | None, pooled (B,L,Cp) | None).
    Cp = C+1 rounded up to a multiple of `pad_to`; the extra columns are zeros (a GEMM operand
    with TMA-aligned rows; `linear` accepts the wider input as is)."""
    if x.dtype != torch.float32:
        raise TypeError("fillbit_pool expects the fp32 embeddings the collate delivers")
    if not x.is_contiguous():
        x = x.contiguous()
    B, SL, C = x.shape
    Lr = SL // S
    Cp = -(-(C + 1) // pad_to) * pad_to
    bit = torch.empty((B, SL), dtype=torch.float32, device=x.device) if want_bit else None
    cat = torch.empty((B, SL, C + 1), dtype=torch.float32, device=x.device) if want_cat else None
    pooled = None
    if want_pooled:
        pooled = torch.empty((B, Lr, Cp), dtype=pooled_dtype or _compute_dtype, device=x.device)
    L.call("dl_fillbit_pool", x.data_ptr(), L.ptr(bit), L.ptr(cat), L.ptr(pooled),
           L.dt(pooled) if pooled is not None else 0, B, S, Lr, C, Cp)
    return bit, cat, pooled


def expand_rows(rows: torch.Tensor, offsets: torch.Tensor, out: torch.Tensor, repeat: bool) -> torch.Tensor:
    """Packed per-sample rows [sum R_b, C] + offsets [B+1] (int32) -> dense out (B, maxsize, C) fp32,
    exactly as utils.tail_pad (repeat=False) / utils.repeat_pad (repeat=True) build it on the host."""
    if rows.dtype != torch.float32 or out.dtype != torch.float32 or offsets.dtype != torch.int32:
        raise TypeError("expand_rows: rows/out must be fp32 and offsets int32")
    B, maxsize, C = out.shape
    if offsets.numel() != B + 1 or rows.shape[-1] != C or not out.is_contiguous() or not rows.is_contiguous():
        raise ValueError("expand_rows: inconsistent shapes")
    L.call("dl_expand_rows", rows.data_ptr(), offsets.data_ptr(), out.data_ptr(), B, maxsize, C, int(repeat))
    return out


def _tok_dtype(tokens: torch.Tensor) -> int:
    if tokens.dtype == torch.int64:
        return 0
    if tokens.dtype == torch.float64:
        return 1
    raise TypeError(f"tokens must be int64 or float64 (got {tokens.dtype})")


def embed_fill_fwd(tokens: torch.Tensor, fill: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """tokens (B, L) int64/float64, fill (B, L) fp32, table (V, 127) fp32 -> (B, L, 128) compute dtype."""
    B, Ls = tokens.shape
    out = torch.empty((B, Ls, table.shape[1] + 1), dtype=_compute_dtype, device=table.device)
    L.call("dl_embed_fill_fwd", tokens.data_ptr(), _tok_dtype(tokens), fill.data_ptr(), table.data_ptr(),
           out.data_ptr(), B * Ls, table.shape[0], table.shape[1] + 1, L.dt(out))
    return out


def embed_fill_bwd(tokens: torch.Tensor, g: torch.Tensor, dtable: torch.Tensor, padding_idx: int) -> None:
    """dtable (V, 127) fp32 += per-token sums of g (B, L, 128)[..., :127]."""
    L.call("dl_embed_fill_bwd", tokens.data_ptr(), _tok_dtype(tokens), g.data_ptr(), dtable.data_ptr(),
           tokens.numel(), dtable.shape[0], g.shape[-1], padding_idx, L.dt(g))


def transpose_last2(x: torch.Tensor) -> torch.Tensor:
    """(B, R, C) -> (B, C, R), contiguous."""
    B, R, Cc = x.shape
    y = torch.empty((B, Cc, R), dtype=x.dtype, device=x.device)
    L.call("dl_transpose", x.data_ptr(), y.data_ptr(), B, R, Cc, L.dt(x))
    return y


def conv1d_same(x: torch.Tensor, w_taps: torch.Tensor, out: torch.Tensor, *, taps: int, left: int,
                bias=None, act: int = ACT_NONE) -> torch.Tensor:
    """Channels-last 'same' conv1d as an implicit GEMM.  x (B, L, Cin), w_taps (Cout, taps*Cin) with
    w_taps[co, t*Cin + ci] = weight of tap t; out (B, L, Cout)."""
    B, Ls, Cin = x.shape
    Cout = w_taps.shape[0]
    L.gemm(x, w_taps, out, M=Ls, N=Cout, K=taps * Cin, lda=Cin, ldb=taps * Cin, ldc=Cout,
           batch=(1, 1, B), sa=(0, 0, Ls * Cin), sb=(0, 0, 0), sc=(0, 0, Ls * Cout), bias=bias, act=act,
           conv_taps=taps, conv_left=left)
    return out


def conv1d_same_wgrad(g: torch.Tensor, x: torch.Tensor, taps: int, left: int) -> torch.Tensor:
    """dW[t, co, ci] = sum_{b,l} g[b, l, co] * x[b, l + t - left, ci]  (fp32, (taps, Cout, Cin))."""
    B, Ls, Cout = g.shape
    Cin = x.shape[2]
    dw = torch.empty((taps, Cout, Cin), dtype=torch.float32, device=g.device)
    L.gemm(g, x, dw, M=Cout, N=Cin, K=Ls, lda=Cout, ldb=Cin, ldc=Cin, trans_a=True, trans_b=True,
           batch=(taps, 1, B), sa=(0, 0, Ls * Cout), sb=(0, 0, Ls * Cin), sc=(Cout * Cin, 0, 0),
           kred=True, kred_shift=-left)
    return dw


def site_pool_fwd(x: torch.Tensor, S: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (B, S*L, C) -> (B, L, C); `out` may be a column-slice view (row stride = ldy)."""
    B, SL, C = x.shape
    Lr = SL // S
    if out is None:
        out = torch.empty((B, Lr, C), dtype=x.dtype, device=x.device)
    L.call("dl_site_pool_fwd", x.data_ptr(), out.data_ptr(), B, S, Lr, C, out.stride(1), L.dt(x))
    return out


def site_pool_bwd(dy: torch.Tensor, S: int) -> torch.Tensor:
    B, Lr, C = dy.shape
    dx = torch.empty((B, S * Lr, C), dtype=dy.dtype, device=dy.device)
    L.call("dl_site_pool_bwd", dy.data_ptr(), dx.data_ptr(), B, S, Lr, C, dy.stride(1), L.dt(dy))
    return dx


def mhla_gate_ln_fwd(v: torch.Tensor, logits: torch.Tensor, gamma, beta, eps: float):
    """gamma=None -> gating only (no residual / LayerNorm)."""
    B, Lr, E = v.shape
    H = logits.shape[-1]
    y = torch.empty_like(v)
    p = torch.empty((B, H, Lr), dtype=torch.float32, device=v.device)
    mean = rstd = None
    if gamma is not None:
        mean = torch.empty(B * Lr, dtype=torch.float32, device=v.device)
        rstd = torch.empty_like(mean)
    L.call("dl_mhla_gate_ln_fwd", v.data_ptr(), logits.data_ptr(), L.ptr(gamma), L.ptr(beta),
           y.data_ptr(), p.data_ptr(), L.ptr(mean), L.ptr(rstd), B, Lr, E, H, eps, L.dt(v))
    return y, p, mean, rstd


def mhla_gate_ln_bwd(dy, v, p, mean, rstd, gamma):
    B, Lr, E = v.shape
    H = p.shape[1]
    dv = torch.empty_like(v)
    dlogits = torch.empty((B, Lr, H), dtype=v.dtype, device=v.device)
    dg = db = None
    if gamma is not None:
        dg = torch.empty(E, dtype=torch.float32, device=v.device)
        db = torch.empty_like(dg)
    L.call("dl_mhla_gate_ln_bwd", dy.data_ptr(), v.data_ptr(), p.data_ptr(), L.ptr(mean),
           L.ptr(rstd), L.ptr(gamma), dv.data_ptr(), dlogits.data_ptr(), L.ptr(dg),
           L.ptr(db), B, Lr, E, H, L.dt(v))
    return dv, dlogits, dg, db


def cm_triplet_fwd(cos: torch.Tensor, G: torch.Tensor, margin: float):
    P, D = cos.shape
    acc = torch.empty(2, dtype=torch.float64, device=cos.device)
    loss = torch.empty((), dtype=torch.float32, device=cos.device)
    L.call("dl_cm_triplet_fwd", cos.data_ptr(), G.data_ptr(), P, D, margin, acc.data_ptr(), loss.data_ptr())
    return loss, acc


def cm_triplet_bwd(cos, G, margin: float, acc, gout):
    P, D = cos.shape
    dcos = torch.empty_like(cos)
    L.call("dl_cm_triplet_bwd", cos.data_ptr(), G.data_ptr(), P, D, margin, acc.data_ptr(),
           gout.data_ptr(), dcos.data_ptr())
    return dcos


def cross_entropy_fwd(x: torch.Tensor, labels: torch.Tensor, C: int, ignore_index: int = 0,
                      extra=None, wextra=None):
    """x: 2-D (rows, >=C) view; labels int64 (rows,).  Returns (mean loss, acc workspace)."""
    acc = torch.empty(2, dtype=torch.float64, device=x.device)
    loss = torch.empty((), dtype=torch.float32, device=x.device)
    L.call("dl_cross_entropy_fwd", x.data_ptr(), labels.data_ptr(), L.ptr(extra), L.ptr(wextra),
           x.shape[0], C, _ld(x), ignore_index, acc.data_ptr(), loss.data_ptr(), L.dt(x))
    return loss, acc


def cross_entropy_bwd(x, labels, C, acc, gout, ignore_index: int = 0, extra=None, wextra=None):
    dx = torch.zeros_like(x) if x.shape[1] != C else torch.empty_like(x)
    dextra = torch.empty(x.shape[0], dtype=torch.float32, device=x.device) if extra is not None else None
    L.call("dl_cross_entropy_bwd", x.data_ptr(), labels.data_ptr(), L.ptr(extra), L.ptr(wextra),
           x.shape[0], C, _ld(x), ignore_index, acc.data_ptr(), gout.data_ptr(), dx.data_ptr(),
           L.ptr(dextra), L.dt(x))
    return dx, dextra


def bce_fwd(score: torch.Tensor, y: torch.Tensor):
    n = score.numel()
    prob = torch.empty(n, dtype=torch.float32, device=score.device)
    loss = torch.empty((), dtype=torch.float32, device=score.device)
    L.call("dl_bce_fwd", score.data_ptr(), y.data_ptr(), prob.data_ptr(), loss.data_ptr(), n)
    return prob, loss


def bce_bwd(prob, y, gout):
    ds = torch.empty_like(prob)
    L.call("dl_bce_bwd", prob.data_ptr(), y.data_ptr(), gout.data_ptr(), ds.data_ptr(), prob.numel())
    return ds


# ---- small-M dense layers (the decoder head) ---------------------------------------------------
SMALL_M = 64


def small_linear(x: torch.Tensor, w: torch.Tensor, bias=None, *, w_kn: bool = False, act: int = ACT_NONE,
                 keep_pre: bool = False, bn=None):
    """y = BatchNorm(act(x op(w) + bias)) for at most SMALL_M rows, fp32, one launch (dl_small_linear).
    w: [N, K] (w_kn=False) or [K, N] (w_kn=True), unit inner stride.
    bn: None or (gamma, beta, running_mean, running_var, num_batches_tracked, eps, momentum, training).
    -> (y, pre or None, mean or None, rstd or None)"""
    M, Kd = x.shape
    N = w.shape[1] if w_kn else w.shape[0]
    if x.dtype != torch.float32 or w.dtype != torch.float32 or x.stride(1) != 1 or w.stride(1) != 1:
        raise TypeError("dl_small_linear takes fp32 operands with unit inner stride")
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    pre = torch.empty_like(y) if keep_pre else None
    mean = rstd = None
    a = L.SmallLinearArgs()
    a.X, a.W, a.bias, a.pre, a.Y = x.data_ptr(), w.data_ptr(), L.ptr(bias), L.ptr(pre), y.data_ptr()
    a.M, a.N, a.K = M, N, Kd
    a.ldx, a.ldw, a.ldy = x.stride(0), w.stride(0), N
    a.w_kn, a.act = int(w_kn), act
    if bn is not None:
        gamma, beta, rm, rv, nbt, eps, momentum, training = bn
        mean = torch.empty(N, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        a.gamma, a.beta, a.mean, a.rstd = L.ptr(gamma), L.ptr(beta), mean.data_ptr(), rstd.data_ptr()
        a.running_mean, a.running_var, a.num_batches_tracked = L.ptr(rm), L.ptr(rv), L.ptr(nbt)
        a.bn, a.training, a.eps, a.momentum = 1, int(training), eps, momentum
    L.check(L.lib().dl_small_linear(L.C.byref(a), L.stream_ptr()), "dl_small_linear")
    return y, pre, mean, rstd


def head_bn_act_bwd(dy, pre, gamma, mean, rstd, act: int, training: bool, has_bias: bool = True, acc_into=None):
    """BatchNorm1d + activation backward of one head layer (dl_head_bn_act_bwd).
    -> (g = d loss / d pre, dgamma, dbeta, dbias).  acc_into = (dgamma, dbeta, dbias) fp32 gradient
    buffers to ADD to (entries that do not apply are None); the gradients are then returned as None."""
    M, N = dy.shape
    g = torch.empty_like(dy)
    bn = mean is not None
    if acc_into is not None:
        dg, db, dbias = acc_into
    else:
        new = lambda: torch.empty(N, dtype=torch.float32, device=dy.device)       # noqa: E731
        dg, db = (new(), new()) if (bn and gamma is not None) else (None, None)
        dbias = new() if has_bias else None
    L.call("dl_head_bn_act_bwd", dy.data_ptr(), pre.data_ptr(), L.ptr(gamma), L.ptr(mean), L.ptr(rstd),
           g.data_ptr(), L.ptr(dg), L.ptr(db), L.ptr(dbias), M, N, act, int(bn), int(training),
           int(acc_into is not None))
    if acc_into is not None:
        return g, None, None, None
    return g, dg, db, dbias


def copy_rows(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[r, :] = src[r, :] for 2-D tensors with unit inner stride and any row strides (dl_copy_rows)."""
    if src.shape != dst.shape or src.dtype != dst.dtype or src.stride(1) != 1 or dst.stride(1) != 1:
        raise ValueError("copy_rows needs two (rows, cols) tensors of one dtype with unit inner stride")
    es = src.element_size()
    L.call("dl_copy_rows", src.data_ptr(), src.stride(0) * es, dst.data_ptr(), dst.stride(0) * es,
           src.shape[1] * es, src.shape[0])
    return dst
