"""GPU: dl_gemm (tcgen05 + TMA) against a float64 matmul of the same operands.

Integer-valued operands make every product and partial sum exactly representable, so the
layout / descriptor / pipeline logic is checked bit-exactly; random operands check the
numerics within the tolerance of the input precision (bf16 inputs are exact, tf32 drops 13
mantissa bits of each fp32 operand)."""
import itertools

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, dtype, integer, gen):
    if integer:
        t = torch.randint(-3, 4, shape, generator=gen, device="cuda").to(torch.float32)
    else:
        t = torch.randn(shape, generator=gen, device="cuda", dtype=torch.float32)
    return t.to(dtype)


def run_case(M, N, K, dtype, trans_a, trans_b, tile_n=0, integer=True, batch=(1, 1), out_dtype=None,
             bias=False, act=0, residual=False, preact=False, mul_mode=0, alpha=1.0, pad=0, seed=0):
    from druglamp_b200 import _lib
    gen = torch.Generator(device="cuda").manual_seed(seed)
    blo, bhi = batch
    out_dtype = out_dtype or dtype
    lda = (M if trans_a else K) + pad
    ldb = (N if trans_b else K) + pad
    ldc = N + pad
    A = _mk((bhi, blo, K if trans_a else M, lda), dtype, integer, gen)
    B = _mk((bhi, blo, K if trans_b else N, ldb), dtype, integer, gen)
    Cc = torch.full((bhi, blo, M, ldc), float("nan"), device="cuda", dtype=out_dtype)
    bias_t = _mk((N,), torch.float32, integer, gen) if bias else None
    res_t = _mk((bhi, blo, M, ldc), out_dtype, integer, gen) if residual else None
    aux_t = _mk((bhi, blo, M, ldc), out_dtype, False, gen) if mul_mode else None
    pre_t = torch.zeros_like(Cc) if preact else None
    _lib.gemm(A, B, Cc, M=M, N=N, K=K, lda=lda, ldb=ldb, ldc=ldc, trans_a=trans_a, trans_b=trans_b,
              batch_lo=blo, batch_hi=bhi, sa=(A.stride(1), A.stride(0)), sb=(B.stride(1), B.stride(0)),
              sc=(Cc.stride(1), Cc.stride(0)), alpha=alpha, bias=bias_t, act=act, preact_out=pre_t,
              mul_aux=aux_t, mul_mode=mul_mode, residual=res_t, tile_n=tile_n)
    torch.cuda.synchronize()
    Ad = A.double()[..., :M] .transpose(-1, -2) if trans_a else A.double()[..., :K]
    Bd = B.double()[..., :N] if trans_b else B.double()[..., :K].transpose(-1, -2)
    ref = alpha * (Ad @ Bd)
    if bias:
        ref = ref + bias_t.double()
    pre_ref = ref.clone()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    if mul_mode == 3:
        ref = ref * aux_t.double()[..., :N]
    elif mul_mode == 2:
        ref = ref * (aux_t.double()[..., :N] > 0)
    elif mul_mode == 1:
        a = aux_t.double()[..., :N].requires_grad_(True)
        (g,) = torch.autograd.grad(torch.nn.functional.gelu(a).sum(), a)
        ref = ref * g
    if residual:
        ref = ref + res_t.double()[..., :N]
    got = Cc.double()[..., :N]
    assert torch.isfinite(got).all(), "non-finite output (tile not written?)"
    if pad:
        assert torch.isnan(Cc[..., N:]).all(), "wrote outside the N extent"
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-30
    pre_err = None
    if preact:
        pre_err = (pre_t.double()[..., :N] - pre_ref).abs().max().item()
    return err, scale, pre_err


TRANS = list(itertools.product([False, True], [False, True]))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("trans_a,trans_b", TRANS)
@pytest.mark.parametrize("tile_n", [64, 128, 256])
def test_exact_integer_operands(dtype, trans_a, trans_b, tile_n):
    out_dtype = torch.float32
    for (M, N, K) in [(128, tile_n, 64), (256, 256, 256), (200, 136, 328), (128, 8, 1024)]:
        err, scale, _ = run_case(M, N, K, dtype, trans_a, trans_b, tile_n=tile_n, out_dtype=out_dtype)
        assert err == 0.0, (M, N, K, err, scale)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("trans_a,trans_b", TRANS)
def test_batched_strided_and_padded(dtype, trans_a, trans_b):
    err, scale, _ = run_case(256, 192, 128, dtype, trans_a, trans_b, batch=(3, 2), pad=8,
                             out_dtype=torch.float32)
    assert err == 0.0, (err, scale)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-2), (torch.float32, 2e-3)])
def test_random_numerics_and_epilogue(dtype, tol):
    for kw in [dict(), dict(bias=True, act=1, preact=True), dict(bias=True, act=2, residual=True),
               dict(mul_mode=1, alpha=0.5), dict(mul_mode=2), dict(mul_mode=3, residual=True)]:
        err, scale, pre_err = run_case(384, 320, 512, dtype, False, False, integer=False, **kw)
        assert err <= tol * scale, (kw, err, scale)
        if pre_err is not None:
            assert pre_err <= tol * scale, (kw, pre_err, scale)


def test_long_k_pipeline_wraps():
    err, scale, _ = run_case(128, 128, 4096, torch.bfloat16, False, False, out_dtype=torch.float32)
    assert err == 0.0


def test_argument_errors_are_reported():
    from druglamp_b200 import _lib
    A = torch.zeros(16, 75, device="cuda")
    B = torch.zeros(16, 75, device="cuda")
    Cc = torch.zeros(16, 16, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of 16 bytes"):
        _lib.gemm(A, B, Cc, M=16, N=16, K=75, lda=75, ldb=75, ldc=16)
