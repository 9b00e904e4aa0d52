"""GPU: each non-GEMM kernel of libdruglamp_sm100.so against a plain PyTorch fp32/fp64 reference of
the same op (these are the floating-point row kernels; index/mask outputs are checked bit-exactly)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)]


def _close(a, b, tol, what=""):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item()
    scale = b.abs().max().item() + 1e-12
    assert err <= tol * scale, f"{what}: err {err:.3e} scale {scale:.3e}"


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("cols", [128, 256, 512])
def test_layernorm_fwd_bwd(dtype, tol, cols):
    from druglamp_b200 import kernels as K
    torch.manual_seed(0)
    rows = 1000
    x = (torch.randn(rows, cols, device="cuda") * 2 + 0.5).to(dtype)
    g = torch.randn(cols, device="cuda")
    b = torch.randn(cols, device="cuda")
    dy = torch.randn(rows, cols, device="cuda").to(dtype)
    y, mean, rstd = K.layernorm_fwd(x, g, b, 1e-6)
    dx, dg, db = K.layernorm_bwd(dy, x, g, mean, rstd)
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.layer_norm(xr, (cols,), gr, br, 1e-6)
    yr.backward(dy.double())
    _close(y, yr, tol, "y")
    _close(dx, xr.grad, tol, "dx")
    _close(dg, gr.grad, max(tol, 1e-4), "dgamma")
    _close(db, br.grad, max(tol, 1e-4), "dbeta")
    # dx_add: the skip connection's gradient joins inside the kernel (pre-norm residual blocks)
    extra = torch.randn(rows, cols, device="cuda").to(dtype)
    dx2, _, _ = K.layernorm_bwd(dy, x, g, mean, rstd, dx_add=extra)
    _close(dx2, xr.grad + extra.double(), tol, "dx + skip")


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("cols", [256, 290, 512])
def test_softmax_fwd_bwd(dtype, tol, cols):
    from druglamp_b200 import kernels as K
    torch.manual_seed(1)
    s = (torch.randn(77, 3, cols, device="cuda") * 3).to(dtype)
    dp = torch.randn(77, 3, cols, device="cuda").to(dtype)
    p = K.softmax_fwd(s.clone())
    sr = s.double().requires_grad_(True)
    pr = F.softmax(sr, dim=-1)
    _close(p, pr, tol, "p")
    ds = K.softmax_bwd(p, dp.clone(), 0.5)
    (gr,) = torch.autograd.grad(pr, sr, dp.double())
    _close(ds, 0.5 * gr, max(tol, 1e-4) * 2, "ds")


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("cols", [70, 290, 513, 1000])
def test_softmax_padded_rows_leave_the_padding_untouched(dtype, tol, cols):
    """Rows padded to a 16-byte stride (the score maps of a key length such as 290): the ragged
    vector kernels must give the same result and never write the pad columns."""
    from druglamp_b200 import kernels as K
    torch.manual_seed(2)
    ld = (cols + 7) // 8 * 8 + 8                                     # at least one full pad run
    full = (torch.randn(5, 41, ld, device="cuda") * 3).to(dtype)
    full[..., cols:] = 777.0
    s = full[..., :cols]
    dp_full = torch.randn(5, 41, ld, device="cuda").to(dtype)
    dp_full[..., cols:] = 555.0
    out_full = torch.full_like(full, 333.0)
    p = K.softmax_fwd(s, out=out_full[..., :cols])
    sr = s.double().requires_grad_(True)
    pr = F.softmax(sr, dim=-1)
    _close(p, pr, tol, "p")
    assert (out_full[..., cols:] == 333.0).all() and (full[..., cols:] == 777.0).all()
    dpv = dp_full[..., :cols]
    (gr,) = torch.autograd.grad(pr, sr, dpv.double())
    ds = K.softmax_bwd(p, dpv, 0.5)
    _close(ds, 0.5 * gr, max(tol, 1e-4) * 2, "ds")
    assert (dp_full[..., cols:] == 555.0).all() and (out_full[..., cols:] == 333.0).all()


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_colsum_actbwd_dropout_cast_addpe(dtype, tol):
    from druglamp_b200 import kernels as K
    torch.manual_seed(2)
    x = torch.randn(3000, 264, device="cuda").to(dtype)
    _close(K.colsum(x), x.double().sum(0), 1e-4, "colsum")
    _close(K.colsum(x[:, 8:136]), x.double()[:, 8:136].sum(0), 1e-4, "colsum view")
    pre = torch.randn_like(x)
    a = pre.double().requires_grad_(True)
    (gg,) = torch.autograd.grad(F.gelu(a).sum(), a)
    _close(K.act_bwd(x, pre, K.ACT_GELU), x.double() * gg, tol, "gelu bwd")
    _close(K.act_bwd(x, pre, K.ACT_RELU), x.double() * (pre.double() > 0), tol, "relu bwd")
    d1 = K.dropout(x, 0.1, 1234)
    d2 = K.act_bwd(x, None, K.ACT_NONE, (0.1, 1234))
    assert torch.equal(d1, d2)
    keep = (d1 != 0) | (x == 0)
    assert 0.86 < keep.float().mean().item() < 0.94
    _close(d1[keep], (x.double() / 0.9)[keep], tol, "dropout scale")
    y = K.cast(x, torch.float32)
    assert torch.equal(y, x.float())
    z = K.cast(y, torch.bfloat16)
    assert torch.equal(z, y.to(torch.bfloat16))
    odd = torch.randn(1003, device="cuda")
    assert torch.equal(K.cast(odd, torch.bfloat16), odd.to(torch.bfloat16))
    pe = torch.randn(50, 264, device="cuda")
    xb = x.view(60, 50, 264)
    _close(K.add_pe(xb, pe), xb.double() + pe.double(), tol, "add_pe")


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_spmm_norm_matches_index_add(dtype, tol):
    from druglamp_b200 import kernels as K
    from druglamp_b200.synth import make_batch
    b = make_batch(3, seed=5)
    g = b.graph.to("cuda")
    torch.manual_seed(3)
    h = torch.randn(g.num_nodes(), 128, device="cuda").to(dtype)
    out = K.spmm_norm(g.indptr, g.indices, g.norm_src, g.norm_dst, h)
    hd = h.double() * g.norm_src.double()[:, None]
    ref = torch.zeros_like(hd).index_add_(0, g.dst, hd[g.src]) * g.norm_dst.double()[:, None]
    _close(out, ref, tol, "spmm")
    # backward operator = transposed CSR with swapped norms
    outT = K.spmm_norm(g.indptr_t, g.indices_t, g.norm_dst, g.norm_src, h)
    hd = h.double() * g.norm_dst.double()[:, None]
    refT = torch.zeros_like(hd).index_add_(0, g.src, hd[g.dst]) * g.norm_src.double()[:, None]
    _close(outT, refT, tol, "spmm^T")


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_fwd_bwd(dtype, tol, training):
    from druglamp_b200 import kernels as K
    torch.manual_seed(4)
    rows, cols = 5000, 128
    x = (torch.randn(rows, cols, device="cuda") * 1.5 + 0.3).to(dtype)
    dy = torch.randn(rows, cols, device="cuda").to(dtype)
    bn = torch.nn.BatchNorm1d(cols).cuda().double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
        bn.running_mean.uniform_(-0.2, 0.2)
        bn.running_var.uniform_(0.8, 1.2)
    bn.train(training)
    g, b = bn.weight.detach().float(), bn.bias.detach().float()
    rm, rv = bn.running_mean.detach().float().clone(), bn.running_var.detach().float().clone()
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    y, mean, rstd = K.batchnorm_fwd(x, g, b, rm, rv, nbt, 1e-5, 0.1, training)
    dx, dg, db = K.batchnorm_bwd(dy, x, g, mean, rstd, training)
    xr = x.double().requires_grad_(True)
    yr = bn(xr)
    yr.backward(dy.double())
    _close(y, yr, tol, "y")
    _close(dx, xr.grad, tol, "dx")
    _close(dg, bn.weight.grad, max(tol, 1e-4), "dgamma")
    _close(db, bn.bias.grad, max(tol, 1e-4), "dbeta")
    if training:
        _close(rm, bn.running_mean, 1e-4, "running_mean")
        _close(rv, bn.running_var, 1e-4, "running_var")
        assert int(nbt) == 1
    # relu_mask: the same dx with the ReLU mask of the BatchNorm input folded in
    dxm, _, _ = K.batchnorm_bwd(dy, x, g, mean, rstd, training, relu_mask=True)
    assert torch.equal(dxm, torch.where(x > 0, dx, torch.zeros_like(dx)))


def test_fillbit_pool_bit_exact_and_site_pool():
    from druglamp_b200 import kernels as K
    from druglamp_b200.synth import make_batch
    b = make_batch(3, seed=11)
    xp = b.xp.cuda()
    bit, cat, pooled = K.fillbit_pool(xp, 9, want_cat=True, pooled_dtype=torch.float32)
    ref_bit = (xp.sum(-1) == 0).float()
    assert torch.equal(bit, ref_bit)                           # bit-exact padding mask
    ref_cat = torch.cat((xp, ref_bit.unsqueeze(-1)), -1)
    assert torch.equal(cat, ref_cat)
    ref_pool = ref_cat.view(-1, 9, 256, 641).mean(1)
    _close(pooled, ref_pool, 1e-6, "pooled")
    xd = b.xd.cuda()
    bit_d, cat_d, _ = K.fillbit_pool(xd, 1, want_cat=True, want_pooled=False)
    assert torch.equal(bit_d, (xd.sum(-1) == 0).float())
    assert torch.equal(cat_d, torch.cat((xd, bit_d.unsqueeze(-1)), -1))
    for dtype, tol in DTYPES:
        v = torch.randn(3, 2304, 128, device="cuda").to(dtype)
        buf = torch.zeros(3, 256, 256, device="cuda", dtype=dtype)
        K.site_pool_fwd(v, 9, out=buf[:, :, :128])
        _close(buf[:, :, :128], v.double().view(3, 9, 256, 128).mean(1), tol, "site pool")
        assert (buf[:, :, 128:] == 0).all()
        dy = torch.randn(3, 256, 128, device="cuda").to(dtype)
        dx = K.site_pool_bwd(dy, 9)
        _close(dx, (dy.double() / 9)[:, None].expand(3, 9, 256, 128).reshape(3, 2304, 128), tol, "site pool bwd")


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("shape", [(3, 256, 256, 8), (2, 100, 128, 8)])   # H | L: one block per head; else per pair
def test_mhla_gate_ln_fwd_bwd(dtype, tol, shape):
    from druglamp_b200 import kernels as K
    torch.manual_seed(6)
    B, Lr, E, H = shape
    v = torch.randn(B, Lr, E, device="cuda").to(dtype)
    logits = (torch.randn(B, Lr, H, device="cuda") * 2).to(dtype)
    g = torch.rand(E, device="cuda") + 0.5
    bt = torch.randn(E, device="cuda")
    dy = torch.randn(B, Lr, E, device="cuda").to(dtype)
    y, p, mean, rstd = K.mhla_gate_ln_fwd(v, logits, g, bt, 1e-5)
    dv, dlog, dg, db = K.mhla_gate_ln_bwd(dy, v, p, mean, rstd, g)
    vr = v.double().requires_grad_(True)
    lr_ = logits.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), bt.double().requires_grad_(True)
    a = F.softmax(lr_, dim=1).transpose(1, 2).contiguous()              # reference encoder.py:132-140
    gated = (a.view(B * H, Lr, 1) * vr.contiguous().view(B * H, Lr, E // H)).view(B, Lr, E)
    yr = F.layer_norm(gated + vr, (E,), gr, br, 1e-5)
    yr.backward(dy.double())
    _close(y, yr, tol, "y")
    _close(p, a, 1e-5 if dtype == torch.float32 else tol, "p")
    # dv here is only the direct (gate/residual) path == full grad w.r.t. v with logits held fixed
    _close(dv, vr.grad, tol, "dv")
    _close(dlog, lr_.grad, max(tol, 1e-4) * 2, "dlogits")
    _close(dg, gr.grad, max(tol, 1e-4), "dgamma")
    _close(db, br.grad, max(tol, 1e-4), "dbeta")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("k", [3, 6, 9])
def test_conv1d_same_implicit_gemm_fwd_bwd(dtype, tol, k):
    """Conv1d(padding='same') + ReLU as implicit GEMM (TMA row shifts, zero fill) vs F.conv1d."""
    import druglamp_b200 as D
    from druglamp_b200 import functions as Fn
    D.set_compute_dtype(dtype)
    try:
        torch.manual_seed(8 + k)
        B, Ls, Cin, Cout = 3, 384, 128, 128
        x = torch.randn(B, Ls, Cin, device="cuda").to(dtype).float().requires_grad_(True)
        w = (torch.randn(Cout, Cin, k, device="cuda") * 0.05).to(dtype).float().requires_grad_(True)
        b = torch.randn(Cout, device="cuda").requires_grad_(True)
        gy = torch.randn(B, Ls, Cout, device="cuda").to(dtype).float()
        y = Fn.Conv1dSameFn.apply(x, torch.nn.Parameter(w.detach().clone()), torch.nn.Parameter(b.detach().clone()), True, True)
        xr, wr, br = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
        yr = F.relu(F.conv1d(xr.transpose(1, 2), wr, br, padding="same")).transpose(1, 2)
        _close(y, yr, tol, "conv fwd")
        # backward through fresh leaves
        wp, bp = torch.nn.Parameter(w.detach().clone()), torch.nn.Parameter(b.detach().clone())
        xl = x.detach().clone().requires_grad_(True)
        y2 = Fn.Conv1dSameFn.apply(xl, wp, bp, True, True)
        y2.backward(gy.to(y2.dtype))
        # use the product's own ReLU mask so that bf16 rounding at the kink does not enter the comparison
        mask = (y2.detach().double() > 0)
        pre = F.conv1d(xr.transpose(1, 2), wr, br, padding="same").transpose(1, 2)
        (pre * mask * gy.double()).sum().backward()
        _close(xl.grad, xr.grad, max(tol, 1e-4), "conv dx")
        _close(wp.grad, wr.grad, max(tol, 1e-4), "conv dw")
        _close(bp.grad, br.grad, max(tol, 1e-4), "conv db")
        t = torch.randn(2, 70, 45, device="cuda").to(dtype)
        from druglamp_b200 import kernels as K
        assert torch.equal(K.transpose_last2(t), t.transpose(1, 2).contiguous())
    finally:
        D.set_compute_dtype(torch.float32)


def test_cm_triplet_and_bce():
    from druglamp_b200 import kernels as K
    from oracle import restatement as R
    torch.manual_seed(7)
    P, D = 37, 53
    pl = F.normalize(torch.randn(P, 256, dtype=torch.float64), dim=-1)
    dl = F.normalize(torch.randn(D, 256, dtype=torch.float64), dim=-1)
    G = (torch.rand(P, D) < 0.15).long()
    G[3] = 0                                   # a protein without positives
    G[5] = 1                                   # a protein without negatives
    cos = (pl @ dl.t()).requires_grad_(True)
    for margin in (0.5, 0.0187):
        # oracle on the same cosines (dense form pinned against the reference in test_oracle_golden)
        S = torch.sigmoid(cos)
        total, n = 0.0, 0
        for i in range(P):
            pi, ni = (G[i] == 1).nonzero().flatten(), (G[i] == 0).nonzero().flatten()
            if len(pi) and len(ni):
                total = total + F.relu(S[i, ni][None] - S[i, pi][:, None] + margin).sum(); n += len(pi) * len(ni)
            elif len(ni):
                total = total + F.relu(S[i, ni] - torch.sigmoid(torch.tensor(1.0, dtype=torch.float64)) + margin).sum(); n += len(ni)
        ref = total / max(n, 1)
        (gref,) = torch.autograd.grad(ref, cos)
        c32 = cos.detach().float().cuda()
        loss, acc = K.cm_triplet_fwd(c32, G.to(torch.int8).cuda(), margin)
        assert int(acc[1].item()) == n
        assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
        gout = torch.full((), 2.0, device="cuda")
        dcos = K.cm_triplet_bwd(c32, G.to(torch.int8).cuda(), margin, acc, gout)
        _close(dcos.cpu(), 2.0 * gref, 1e-4, "dcos")
    assert abs(R.cm_triplet_dense(pl.float(), dl.float(), G, 0.5).item() - float(K.cm_triplet_fwd(
        (pl @ dl.t()).float().cuda(), G.to(torch.int8).cuda(), 0.5)[0])) < 1e-5
    score = torch.randn(64, 1, device="cuda") * 3
    y = (torch.rand(64, device="cuda") < 0.45).float()
    prob, loss = K.bce_fwd(score, y)
    sr = score.clone().requires_grad_(True)
    n_ref, l_ref = R.binary_cross_entropy(sr, y)
    l_ref.backward()
    _close(prob, n_ref, 1e-6, "prob")
    assert abs(loss.item() - l_ref.item()) < 1e-5
    ds = K.bce_bwd(prob, y, torch.ones((), device="cuda"))
    _close(ds, sr.grad.flatten(), 1e-5, "dscore")


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("tok_dtype", [torch.int64, torch.float64])
def test_embed_fill_fwd_bwd_matches_embedding_cat(dtype, tol, tok_dtype):
    """ProteinCNN input (basic_model.py:171-173): nn.Embedding(27, 127, padding_idx=0) + cat(fill)."""
    import druglamp_b200 as D
    from druglamp_b200 import functions as Fn
    torch.manual_seed(3)
    B, Ls = 5, 1200
    emb = torch.nn.Embedding(27, 127, padding_idx=0).cuda()
    tokens = torch.randint(0, 26, (B, Ls), device="cuda")
    tokens[:, 900:] = 0                                             # padded tail
    fill = (torch.rand(B, Ls, device="cuda") < 0.3).float()
    gy = torch.randn(B, Ls, 128, device="cuda")
    ref = torch.cat((emb(tokens), fill.unsqueeze(-1)), -1)
    ref.backward(gy.to(dtype).float())
    gref = emb.weight.grad.clone()
    emb.weight.grad = None
    D.set_compute_dtype(dtype)
    try:
        out = Fn.EmbedFillFn.apply(tokens.to(tok_dtype), fill, emb.weight, emb.padding_idx)
        assert out.dtype == dtype and out.shape == (B, Ls, 128)
        _close(out, ref.detach(), 4e-3 if dtype == torch.bfloat16 else 0.0, "embed+fill")
        assert torch.equal(out[..., 127].float(), fill)
        out.backward(gy.to(dtype))
    finally:
        D.set_compute_dtype(torch.float32)
    assert emb.weight.grad[0].abs().max().item() == 0.0             # padding row gets no gradient
    _close(emb.weight.grad, gref, 1e-5, "dtable")


def test_packed_collate_expands_bit_exactly_like_the_reference_pads():
    """dl_expand_rows against the restated utils.tail_pad / utils.repeat_pad (:304-324), ragged
    blocks incl. one longer than half of maxsize (tiled once) and a single-row block; and the
    whole packed input path against the dense synthetic batch."""
    from druglamp_b200.collate import pack_rows
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch
    from oracle import restatement as R
    torch.manual_seed(1)
    blocks = [torch.randn(n, 64) for n in (7, 1, 130, 255, 256, 33)]
    for repeat, ref in ((False, R.tail_pad), (True, R.repeat_pad)):
        pk = pack_rows(blocks, 256, repeat).to("cuda")
        assert torch.equal(pk.dense().cpu(), ref(blocks, 256))
    b = make_batch(6, seed=77)
    sb = StaticBatch(b, torch.device("cuda"))
    host = sb.host_copy_packed(b)
    dense_bytes = sum(t.numel() * t.element_size() for t in sb.host_copy(pin=False))
    sb.xp.fill_(-1.0); sb.xd.fill_(-1.0); sb.vp.zero_()
    n = sb.load_from_packed(host)
    torch.cuda.synchronize()
    assert torch.equal(sb.xp.cpu(), b.xp) and torch.equal(sb.xd.cpu(), b.xd) and torch.equal(sb.vp.cpu(), b.vp)
    assert n < dense_bytes / 2


def test_dropout_step_counter_advances_masks_consistently():
    """With a device step counter every dropout launch derives its seed on the device: the GEMM
    epilogue, dl_dropout and the backward's dl_act_bwd agree at one counter value, and a new value
    gives a different mask with nothing changed on the host (what a CUDA-graph replay sees)."""
    from druglamp_b200 import _lib, kernels as K
    torch.manual_seed(2)
    x = torch.randn(256, 128, device="cuda").to(torch.bfloat16)
    w = torch.eye(128, device="cuda").to(torch.bfloat16)
    ctr = torch.zeros((), dtype=torch.int64, device="cuda")
    K.set_dropout_step(ctr)
    try:
        masks = []
        for step in (0, 1, 1, 2):
            ctr.fill_(step)
            y = torch.empty_like(x)
            _lib.gemm(x, w, y, M=256, N=128, K=128, lda=128, ldb=128, ldc=128, drop_p=0.25, drop_seed=77)
            d = K.dropout(x, 0.25, 77)
            g = K.act_bwd(x, None, K.ACT_NONE, (0.25, 77))
            assert torch.equal(y, d) and torch.equal(d, g)
            masks.append(d != 0)
        assert torch.equal(masks[1], masks[2])
        assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[1], masks[3])
        keep = torch.stack(masks).float().mean().item()
        assert abs(keep - 0.75) < 0.01
    finally:
        K.set_dropout_step(None)
    assert torch.equal(K.dropout(x, 0.25, 77) != 0, K.dropout(x, 0.25, 77) != 0)


def test_csr_build_on_device_equals_host_construction():
    """dl_csr_build (degree count -> scan -> stable scatter) against BatchedMolGraph's host
    construction (torch argsort(stable) + bincount): indptr / indices of both directions and the
    clamp(deg, 1)^-1/2 norms bit for bit, duplicate edges (the double self loops of App. A5) and
    isolated nodes included; rebuild_ refills the same buffers for a second edge list."""
    from druglamp_b200.graph import BatchedMolGraph
    from druglamp_b200.synth import make_batch
    b = make_batch(6, seed=17)
    for src, dst, n in ((b.graph.src, b.graph.dst, b.graph.num_nodes()),
                        (torch.tensor([0, 0, 3, 3, 3, 2, 5, 5]), torch.tensor([1, 1, 0, 3, 3, 2, 0, 4]), 7)):
        host = BatchedMolGraph(src, dst, n, 1)
        dev = BatchedMolGraph.from_edges_device(src.cuda(), dst.cuda(), n, 1)
        for k in ("indptr", "indices", "indptr_t", "indices_t"):
            assert torch.equal(getattr(dev, k).cpu(), getattr(host, k)), k
        for k in ("norm_src", "norm_dst"):             # 1/sqrt(deg): the host pow(-0.5) may differ in the last bit
            assert torch.allclose(getattr(dev, k).cpu(), getattr(host, k), rtol=2e-7, atol=0), k
        assert torch.equal(dev.in_deg.cpu(), host.in_deg) and torch.equal(dev.out_deg.cpu(), host.out_deg)
        dev.check_no_zero_in_degree_quiet()
        assert dev._zero_in == host._zero_in
    # same buffers, new edges (a permutation of the edge list changes the stable order inside rows)
    perm = torch.randperm(b.graph.src.numel(), generator=torch.Generator().manual_seed(1))
    s2, d2 = b.graph.src[perm], b.graph.dst[perm]
    dev = BatchedMolGraph.from_edges_device(b.graph.src.cuda(), b.graph.dst.cuda(), b.graph.num_nodes(), 6)
    ptr0 = dev.indices.data_ptr()
    dev.rebuild_(s2.cuda(), d2.cuda())
    host = BatchedMolGraph(s2, d2, b.graph.num_nodes(), 6)
    assert dev.indices.data_ptr() == ptr0
    for k in ("indptr", "indices", "indptr_t", "indices_t"):
        assert torch.equal(getattr(dev, k).cpu(), getattr(host, k)), k


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_batchnorm_weighted_last_row_equals_expanded_batch(dtype, tol):
    """dl_batchnorm last_row_weight = w: the last row stands for w identical rows (the molecular GCN's
    virtual nodes).  Forward statistics, running buffers, dx, dgamma and dbeta must equal
    nn.BatchNorm1d on the batch with that row written out w times -- the last row's dy being the sum
    of the copies' gradients and its dx the sum of their input gradients."""
    from druglamp_b200 import kernels as K
    torch.manual_seed(9)
    R, w, cols = 700, 333, 128
    x = (torch.randn(R + 1, cols, device="cuda") * 0.7 + 2.0).to(dtype)       # |mean| >> std like the GCN
    dy_real = torch.randn(R, cols, device="cuda")
    dy_copies = torch.randn(w, cols, device="cuda")
    dy = torch.cat((dy_real, dy_copies.sum(0, keepdim=True))).to(dtype)
    bn = torch.nn.BatchNorm1d(cols).cuda().double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    g, b = bn.weight.detach().float(), bn.bias.detach().float()
    rm, rv = bn.running_mean.detach().float().clone(), bn.running_var.detach().float().clone()
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    y, mean, rstd = K.batchnorm_fwd(x, g, b, rm, rv, nbt, 1e-5, 0.1, True, last_row_weight=float(w))
    dx, dg, db = K.batchnorm_bwd(dy, x, g, mean, rstd, True, last_row_weight=float(w))
    xe = torch.cat((x[:R], x[R:].expand(w, cols))).double().requires_grad_(True)
    ye = bn(xe)
    dye = torch.cat((dy[:R].double(), dy_copies.double()))
    ye.backward(dye)
    _close(y[:R], ye[:R], tol, "y real")
    _close(y[R], ye[R], tol, "y virtual")
    _close(dx[:R], xe.grad[:R], max(tol, 2e-5), "dx real")
    _close(dx[R], xe.grad[R:].sum(0), max(tol, 5e-5), "dx virtual (summed)")
    _close(dg, bn.weight.grad, max(tol, 1e-4), "dgamma")
    _close(db, bn.bias.grad, max(tol, 1e-4), "dbeta")
    _close(rm, bn.running_mean, 1e-4, "running_mean")
    _close(rv, bn.running_var, 1e-4, "running_var")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 2e-2)])
def test_gcn_virtual_node_dedup_equals_dense_evaluation(dtype, tol):
    """MolecularGCN on real nodes + one representative virtual node (graph.compact, weighted BatchNorm,
    expansion) against the dense evaluation of all B*512 rows: outputs and every parameter gradient."""
    import druglamp_b200 as D
    from druglamp_b200 import modules as M
    from druglamp_b200.synth import make_batch
    D.set_compute_dtype(dtype)
    try:
        b = make_batch(6, seed=31)
        c = b.graph.compact()
        assert c is not None and c.n_virtual + c.real_idx.numel() == 6 * 512
        res = {}
        for dedup in (False, True):
            M.GCN_DEDUP = dedup
            torch.manual_seed(2)
            m = M.MolecularGCN(75, 128, True, [128] * 3, None).cuda().train()
            g = b.graph.to("cuda")
            g.ndata["h"] = b.graph.ndata["h"].cuda()
            out = m(g)
            gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).cuda().to(out.dtype)
            out.backward(gy)
            res[dedup] = (out.detach().float(), {k: p.grad.clone() for k, p in m.named_parameters()},
                          m.gnn.gnn_layers[2].bn_layer.running_var.clone())
        _close(res[True][0], res[False][0], tol, "vd")
        _close(res[True][2], res[False][2], 1e-3, "running_var")
        gtol = 1e-3 if dtype == torch.float32 else 2e-2      # bf16 mode: the incoming gradient is bf16
        for k, gd in res[False][1].items():
            _close(res[True][1][k], gd, gtol, k)
    finally:
        M.GCN_DEDUP = True
        D.set_compute_dtype(torch.float32)


# ---------------------------------------------------------------------------- small-M head kernels
@pytest.mark.parametrize("M,N,Kd", [(64, 1024, 512), (64, 256, 1024), (16, 1024, 1024), (2, 100, 250), (64, 1, 256),
                                    (7, 8, 1), (33, 1030, 130)])
def test_small_linear_matches_fp64(M, N, Kd):
    """dl_small_linear: y = x W^T + b (and the dX orientation y = g W) on <= 64 rows, fp32 FMA."""
    from druglamp_b200 import kernels as K
    torch.manual_seed(M * 1000 + N)
    x = torch.randn(M, Kd, device="cuda")
    w = torch.randn(N, Kd, device="cuda") / Kd ** 0.5
    b = torch.randn(N, device="cuda")
    y, pre, _, _ = K.small_linear(x, w, b, act=K.ACT_GELU, keep_pre=True)
    ref_pre = x.double() @ w.double().t() + b.double()
    _close(pre, ref_pre, 2e-6, "pre")
    _close(y, F.gelu(ref_pre), 2e-6, "gelu")
    g = torch.randn(M, N, device="cuda")
    dx, _, _, _ = K.small_linear(g, w, None, w_kn=True)
    _close(dx, g.double() @ w.double(), 2e-6, "dx")


@pytest.mark.parametrize("M", [64, 16, 2])
@pytest.mark.parametrize("training", [True, False])
def test_head_layer_matches_torch_linear_gelu_batchnorm(M, training):
    """Fn.head_layer = nn.Linear -> GELU -> nn.BatchNorm1d: outputs, running buffers, every gradient."""
    from druglamp_b200 import functions as Fn
    from druglamp_b200 import kernels as K
    torch.manual_seed(5 + M)
    Kd, N = 512, 1024
    fc, bn = torch.nn.Linear(Kd, N).cuda(), torch.nn.BatchNorm1d(N).cuda()
    fr, br = torch.nn.Linear(Kd, N).cuda().double(), torch.nn.BatchNorm1d(N).cuda().double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
        fr.load_state_dict(fc.state_dict()); br.load_state_dict(bn.state_dict())
    bn.train(training); br.train(training)
    x = torch.randn(M, Kd, device="cuda", requires_grad=True)
    xr = x.detach().double().requires_grad_(True)
    dy = torch.randn(M, N, device="cuda")
    y = Fn.head_layer(x, fc, K.ACT_GELU, bn)
    y.backward(dy)
    yr = br(F.gelu(fr(xr)))
    yr.backward(dy.double())
    tol = 2e-4 if (training and M == 2) else 2e-5          # two-sample statistics amplify rounding
    _close(y, yr, tol, "y")
    _close(x.grad, xr.grad, tol * 5, "dx")
    _close(fc.weight.grad, fr.weight.grad, 2e-3, "dW")      # dl_gemm TF32 / 3xTF32 product over the rows
    _close(fc.bias.grad, fr.bias.grad, tol * 5, "db")
    _close(bn.weight.grad, br.weight.grad, tol * 5, "dgamma")
    _close(bn.bias.grad, br.bias.grad, tol * 5, "dbeta")
    _close(bn.running_mean, br.running_mean, 1e-5, "running_mean")
    _close(bn.running_var, br.running_var, 1e-5, "running_var")
    assert int(bn.num_batches_tracked) == int(br.num_batches_tracked)


def test_decoder_head_small_kernels_equal_the_gemm_path():
    """modules.MLP on 64 rows: the one-launch-per-layer path against the dl_gemm + dl_batchnorm path."""
    from druglamp_b200 import modules as Mo
    torch.manual_seed(11)
    a, b = Mo.MLP(512, 1024, 256, 1).cuda(), Mo.MLP(512, 1024, 256, 1).cuda()
    b.load_state_dict(a.state_dict())
    x = torch.randn(64, 512, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = a(xa)
    prev, Mo.HEAD_SMALL_KERNELS = Mo.HEAD_SMALL_KERNELS, False
    try:
        yb = b(xb)
    finally:
        Mo.HEAD_SMALL_KERNELS = prev
    r = torch.randn(64, 1, device="cuda")     # (a plain sum has zero gradient through train-mode BatchNorm)
    (ya * r).sum().backward(); (yb * r).sum().backward()
    _close(ya, yb, 5e-3, "score")           # the GEMM path multiplies in TF32
    _close(xa.grad, xb.grad, 2e-2, "dx")
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        _close(p.grad, q.grad, 2e-2, n)
    for (n, p), (_, q) in zip(a.named_buffers(), b.named_buffers()):
        _close(p, q, 5e-3, n)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,R,C", [(3, 2304, 128), (2, 72, 40), (1, 64, 64)])
def test_bn_transpose_and_site_pool_view_bwd(dtype, B, R, C):
    """dl_bn_transpose (BatchNorm affine + (B,L,C)->(B,C,L)) and dl_site_pool_view_bwd (the gradient through
    transpose -> .view(B,L,C) -> .view(B,S,L/S,C).mean(1), model/basic_model.py:179 + model/DrugLAMP.py:35-37)
    against plain PyTorch on the same values."""
    from druglamp_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + R + C)
    x = torch.randn(B, R, C, generator=g, device="cuda").to(dtype)
    mean = torch.randn(C, generator=g, device="cuda") * 0.2
    rstd = torch.rand(C, generator=g, device="cuda") + 0.5
    gamma = torch.randn(C, generator=g, device="cuda")
    beta = torch.randn(C, generator=g, device="cuda")
    assert K.bn_transpose_ok(x)
    y = K.bn_transpose(x)
    assert torch.equal(y, x.transpose(1, 2).contiguous())
    y = K.bn_transpose(x, mean, rstd, gamma, beta)
    ref = ((x.float() - mean) * rstd * gamma + beta).transpose(1, 2)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert (y.float() - ref).abs().max() <= tol * ref.abs().max()
    for S in (9, 4, 1):
        if R % S:
            continue
        xr = x.float().clone().requires_grad_(True)
        pooled = xr.transpose(1, 2).contiguous().view(B, R, C).view(B, S, R // S, C).mean(1)
        gp = torch.randn(pooled.shape, generator=g, device="cuda").to(dtype)
        pooled.backward(gp.float())
        dx = K.site_pool_view_bwd(gp, S)
        assert dx.shape == x.shape
        assert (dx.float() - xr.grad).abs().max() <= (1e-6 if dtype == torch.float32 else 1e-2) * xr.grad.abs().max()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("site_len", [0, 9])
def test_protein_cnn_fused_tail_equals_separate_passes(dtype, site_len):
    """ProteinCNN with the fused tail (statistics -> normalise-and-transpose -> site mean, gradient mapped
    straight back) against the separate BatchNorm / transpose / site-pool passes: outputs, BN buffers and
    every parameter gradient."""
    import druglamp_b200 as D
    from druglamp_b200 import modules as M
    D.set_compute_dtype(dtype)
    try:
        torch.manual_seed(3)
        B, Lr = 3, 2304
        tok = torch.randint(0, 27, (B, Lr), device="cuda")
        fill = (tok == 0).float()
        res = []
        for fused in (True, False):
            M.CNN_TAIL_FUSED = fused
            torch.manual_seed(4)
            cnn = M.ProteinCNN(128, [128, 128, 128], [3, 6, 9]).cuda().train(True)
            out = cnn(tok, fill, site_len=site_len)
            w = torch.cos(torch.arange(out.numel(), device="cuda", dtype=torch.float32)).view_as(out)
            (out.float() * w).sum().backward()
            torch.cuda.synchronize()
            res.append((out.detach().float(), {k: p.grad.float().clone() for k, p in cnn.named_parameters()},
                        {k: v.clone() for k, v in cnn.state_dict().items() if "running" in k}))
        (o1, g1, s1), (o2, g2, s2) = res
        tol = 1e-5 if dtype == torch.float32 else 2e-2
        assert o1.shape == (B, Lr // site_len if site_len else Lr, 128)
        assert (o1 - o2).abs().max() <= tol * o2.abs().max()
        for k in s1:
            assert (s1[k] - s2[k]).abs().max() <= 1e-5 * s2[k].abs().max() + 1e-7, k
        for k in g1:
            a, b = g1[k].double().flatten(), g2[k].double().flatten()
            cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))
            assert cos >= (0.99999 if dtype == torch.float32 else 0.995), (k, cos)
    finally:
        M.CNN_TAIL_FUSED = True
        D.set_compute_dtype(torch.float32)


@pytest.mark.parametrize("M,N,Kc", [(16384, 1024, 8), (333, 64, 5), (1000, 2048, 16), (7, 8, 1)])
def test_smallk_mul(M, N, Kc):
    """dl_smallk_mul: (g [M, K] @ w [K, N]) * aux for K <= 16 -- the input gradient of MHLA's 8-head lin2 times
    the stored GELU derivative (model/PMMA/encoder.py:128-131) -- against float64, and against dl_gemm."""
    from druglamp_b200 import kernels as K
    gen = torch.Generator(device="cuda").manual_seed(M + N + Kc)
    ld = 8 if Kc <= 8 else 16
    gbuf = torch.full((M, ld), float("nan"), device="cuda", dtype=torch.bfloat16)     # padding columns: NaN
    gbuf[:, :Kc] = torch.randn(M, Kc, generator=gen, device="cuda").bfloat16()
    g = gbuf[:, :Kc]
    w = torch.randn(Kc, N, generator=gen, device="cuda").bfloat16()
    aux = torch.randn(M, N, generator=gen, device="cuda").bfloat16()
    assert K.smallk_mul_ok(g, w, aux)
    out = K.smallk_mul(g, w, aux)
    torch.cuda.synchronize()
    ref = (g.double() @ w.double()) * aux.double()
    assert torch.isfinite(out.float()).all()
    assert (out.double() - ref).abs().max() <= 1e-2 * ref.abs().max()
    if Kc == 8:
        out2 = K.mm(g.contiguous(), w, tb=True, mul_aux=aux, mul_mode=K.MUL_VALUE)
        assert (out.float() - out2.float()).abs().max() <= 2 ** -7 * out2.float().abs().max()
