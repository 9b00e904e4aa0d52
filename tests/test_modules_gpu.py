"""GPU: the drop-in classes of SURVEY section 8b used on their OWN (not through DrugLAMP.forward), against
the oracle's functional restatement with the same state_dict.  Covers the shapes the full-model
fixtures never reach: BASELINE.json configs[3] / SURVEY 8d "config 4" -- GuidedCrossAttention(128, 1)
with query (1200, N, 128) and key = value (290, N, 128), a key length whose rows are not a multiple of
16 bytes -- the same with S padded to 512, distinct key / value tensors, and MultiHeadLinearAttention /
PairedMultimodelAttention called directly.

Tolerances: fp32 mode 1e-3 of the tensor scale (bar of BASELINE.json north_star), bf16 2e-2 forward;
bf16 gradients 6e-2 of the gradient scale (two bf16 roundings per GEMM on the backward chain)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _check_param_grads(module, ref, gtol):
    """Every parameter gradient within gtol of its own scale; gradients that are zero in exact
    arithmetic (a bias in front of a softmax is shift-invariant) are rounding noise on both sides
    (a sum of terms of the size of the other gradients that cancels only up to the rounding of each
    term: 2^-9 per bf16 term) and get a fraction of the largest gradient of the module as scale."""
    floor = 1e-3 if gtol <= 2e-3 else 5e-2
    gmax = max(float(ref[n].grad.abs().max()) for n, _ in module.named_parameters() if ref[n].grad is not None)
    for n, p in module.named_parameters():
        if ref[n].grad is None:                   # dead in the reference too (embed.py:50-51)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        a, b = p.grad.detach().double().cpu(), ref[n].grad
        scale = max(float(b.abs().max()), floor * gmax)
        err = float((a - b).abs().max()) / scale
        assert err <= gtol, (n, err)


def _sd(module, prefix=""):
    return {prefix + k: v.detach().double().cpu() for k, v in module.state_dict().items()}


def _randomise(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            r = torch.randn(p.shape, generator=g)
            if p.dim() == 1:                      # biases / norm weights: away from the 0 / 1 init
                p.copy_((r * 0.1 + (1.0 if "norm" in n and n.endswith("weight") else 0.0)).to(p))
            elif n.split(".")[-1].startswith("pe_"):
                p.copy_((r * 0.5).to(p))
            else:
                p.copy_((r * p.shape[-1] ** -0.5).to(p))


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float32, 1e-3, 1e-3), (torch.bfloat16, 2e-2, 6e-2)])
@pytest.mark.parametrize("Lq,S,shared", [(1200, 290, True), (1200, 512, True), (256, 290, False), (77, 33, True)])
def test_guided_cross_attention_standalone(dtype, tol, gtol, Lq, S, shared):
    import druglamp_b200 as D
    from druglamp_b200.modules import GuidedCrossAttention
    from oracle import restatement as R
    N, E = 3, 128
    D.set_compute_dtype(dtype)
    try:
        m = GuidedCrossAttention(E, 1).cuda()
        _randomise(m, 5)
        g = torch.Generator().manual_seed(Lq * 1000 + S)
        q0 = torch.randn(Lq, N, E, generator=g)
        k0 = torch.randn(S, N, E, generator=g)
        v0 = k0 if shared else torch.randn(S, N, E, generator=g)
        go = torch.randn(Lq, N, E, generator=g)

        q = q0.cuda().requires_grad_(True)
        k = k0.cuda().requires_grad_(True)
        v = k if shared else v0.cuda().requires_grad_(True)
        out, raw = m(q, k, v)
        assert out.shape == (Lq, N, E) and raw.shape == (N, 1, Lq, S) and raw.is_contiguous()
        out.backward(go.cuda())

        sd = _sd(m)
        qr = q0.double().requires_grad_(True)
        kr = k0.double().requires_grad_(True)
        vr = kr if shared else v0.double().requires_grad_(True)
        pr = {n: sd[n].requires_grad_(True) for n in sd}
        o_ref, raw_ref = R.pgca(pr, "", qr, kr, vr, 1)
        o_ref.backward(go.double())

        assert _rel(out, o_ref) <= tol, ("out", _rel(out, o_ref))
        assert _rel(raw, raw_ref) <= tol, ("raw", _rel(raw, raw_ref))
        assert _rel(q.grad, qr.grad) <= gtol, ("dquery", _rel(q.grad, qr.grad))
        assert _rel(k.grad, kr.grad) <= gtol, ("dkey", _rel(k.grad, kr.grad))
        if not shared:
            assert _rel(v.grad, vr.grad) <= gtol, ("dvalue", _rel(v.grad, vr.grad))
        _check_param_grads(m, pr, gtol)
        # need_weights=False keeps the output and drops the map
        o2, none = m(q, k, v, need_weights=False)
        assert none is None and torch.equal(o2, out)
    finally:
        D.set_compute_dtype(torch.bfloat16)


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float32, 1e-3, 1e-3), (torch.bfloat16, 2e-2, 6e-2)])
@pytest.mark.parametrize("B,Lr,E,H", [(3, 256, 256, 8), (2, 512, 128, 8), (2, 96, 256, 8)])
def test_multi_head_linear_attention_standalone(dtype, tol, gtol, B, Lr, E, H):
    import druglamp_b200 as D
    from druglamp_b200.modules import MultiHeadLinearAttention
    from oracle import restatement as R
    D.set_compute_dtype(dtype)
    try:
        m = MultiHeadLinearAttention(E, H, d_diff=4 * E, dropout=0.0, activation='gelu').cuda()
        _randomise(m, 9)
        g = torch.Generator().manual_seed(B * 100 + Lr)
        v0 = torch.randn(B, Lr, E, generator=g)
        go = torch.randn(B, Lr, E, generator=g)
        v = v0.cuda().requires_grad_(True)
        y = m(v)
        y.backward(go.cuda())
        sd = _sd(m)
        pr = {n: sd[n].requires_grad_(True) for n in sd}
        vr = v0.double().requires_grad_(True)
        y_ref = R.mhla(pr, "", vr)
        y_ref.backward(go.double())
        assert _rel(y, y_ref) <= tol, ("y", _rel(y, y_ref))
        assert _rel(v.grad, vr.grad) <= gtol, ("dv", _rel(v.grad, vr.grad))
        _check_param_grads(m, pr, gtol)
    finally:
        D.set_compute_dtype(torch.bfloat16)


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float32, 1e-3, 2e-3), (torch.bfloat16, 2e-2, 8e-2)])
def test_paired_multimodel_attention_standalone_eval(dtype, tol, gtol):
    """PairedMultimodelAttention(config, vis=False) on its own, dropout off (eval): returns
    (encoded (B,256,512), [], []) like paired_multi_model_attention_model.py:22-29 / encoder.py:41-56."""
    import druglamp_b200 as D
    from druglamp_b200.config import get_model_defaults
    from druglamp_b200.modules import PairedMultimodelAttention
    from oracle import restatement as R
    D.set_compute_dtype(dtype)
    try:
        cfg = get_model_defaults(128)
        m = PairedMultimodelAttention(cfg, vis=False).cuda()
        _randomise(m, 21)
        m.eval()
        B = 3
        g = torch.Generator().manual_seed(4)
        p0 = torch.randn(B, 256, 256, generator=g)
        m0 = torch.randn(B, 256, 256, generator=g)
        go = torch.randn(B, 256, 512, generator=g)
        pc = p0.cuda().requires_grad_(True)
        mc = m0.cuda().requires_grad_(True)
        enc, w1, w2 = m(pc, mc)
        assert enc.shape == (B, 256, 512) and w1 == [] and w2 == []
        enc.backward(go.cuda())
        sd = _sd(m)
        pr = {n: (t.requires_grad_(True) if t.is_floating_point() else t) for n, t in sd.items()}
        p_r = p0.double().requires_grad_(True)
        m_r = m0.double().requires_grad_(True)
        e_ref = R.pmma(pr, "", p_r, m_r, 4)
        e_ref.backward(go.double())
        assert _rel(enc, e_ref) <= tol, ("encoded", _rel(enc, e_ref))
        assert _rel(pc.grad, p_r.grad) <= gtol, ("dprot", _rel(pc.grad, p_r.grad))
        assert _rel(mc.grad, m_r.grad) <= gtol, ("dmol", _rel(mc.grad, m_r.grad))
        _check_param_grads(m, pr, gtol)
    finally:
        D.set_compute_dtype(torch.bfloat16)
