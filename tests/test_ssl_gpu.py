"""GPU: the self-supervised heads (H13) against the fixture generated from the unmodified reference
(tests/golden/make_golden.py::run_ssl): protein MLM through the shared ProteinCNN + LLM-logit head and
drug SimSiam, losses and every SSL parameter gradient.  The MLM mask is sampled on the CPU with the
fixture's torch seed (the sampler itself is pinned bit-exactly in test_abi_and_host_cpu.py), so both
sides mask the same positions."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, digest, N_SAMPLES

pytestmark = pytest.mark.gpu


def _digest_err(actual, gold, floor=0.0):
    a = digest(actual.float())
    k = min(N_SAMPLES, actual.numel())
    scale = max(np.abs(gold[:k]).max(), gold[-1] / max(actual.numel(), 1), floor, 1e-30)
    return float(np.abs(a[:k] - gold[:k]).max() / scale)


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float32, 1e-3, 2e-2), (torch.bfloat16, 3e-2, None)])
def test_ssl_losses_and_grads_match_reference_fixture(dtype, tol, gtol):
    import druglamp_b200 as D
    from druglamp_b200 import models
    from druglamp_b200.ssl import sample_mlm_mask
    from druglamp_b200.synth import make_batch
    from oracle import restatement as R
    fx = load_golden("druglamp_train_b8_ssl.npz")
    B, seed = int(fx["meta_B"]), int(fx["meta_seed"])
    D.set_compute_dtype(dtype)
    try:
        m = models.DrugLAMP(384, 640).cuda()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.train(True)
        b = make_batch(B, seed=seed)
        bc = b.to("cuda")

        def det_load():
            shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
            m.load_state_dict(R.deterministic_state(shapes), strict=True)

        det_load()
        from druglamp_b200.modules import binary_cross_entropy
        outs = m(*bc.model_inputs())
        ssl = outs[2]
        # the fixture was recorded after the classification backward of the same model without a
        # zero_grad in between (make_golden.run_model -> run_ssl), so the shared ProteinCNN's
        # gradients are the sum of both losses: follow the same sequence
        binary_cross_entropy(outs[4], bc.y)[1].backward()
        inp = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in ssl.items()}
        m.ssl_model(**inp)                      # materialises the lazy SimSiam projectors (App. A13)
        assert any(".projector." in k for k in m.state_dict())
        det_load()                              # like the fixture: deterministic weights incl. projectors
        torch.manual_seed(12)
        m.ssl_model._mask_override = sample_mlm_mask(b.vp)          # CPU sampler, fixture's seed
        assert np.array_equal(m.ssl_model._mask_override[0].numpy().astype(np.int8), fx["ssl_labels"])
        out = m.ssl_model(**inp)
        prot, drug = float(out["prot_ssl"]), float(out["drug_ssl"])
        assert abs(prot - float(fx["ssl_prot"])) <= tol * abs(float(fx["ssl_prot"])), (prot, float(fx["ssl_prot"]))
        assert abs(drug - float(fx["ssl_drug"])) <= tol * abs(float(fx["ssl_drug"])), (drug, float(fx["ssl_drug"]))
        (out["prot_ssl"] + out["drug_ssl"]).backward()
        if gtol is not None:
            params = dict(m.ssl_model.named_parameters())
            gmax = max(np.abs(fx[k][:64]).max() for k in fx if k.startswith("ssl_grad/"))
            errs = []
            for k in fx:
                if k.startswith("ssl_grad/"):
                    g = params[k[len("ssl_grad/"):]].grad
                    assert g is not None, k
                    e = _digest_err(g, fx[k], floor=1e-3 * gmax)
                    if k.startswith("ssl_grad/extractor."):
                        e /= 6.0      # ReLU -> BatchNorm mask flips in the ProteinCNN (see test_model_parity_gpu.py)
                    errs.append((e, k, float(np.abs(fx[k][:64]).max()), float(g.abs().max())))
            errs.sort(reverse=True)
            print("gmax %.3e; " % gmax + "; ".join(f"{k[9:]} err={e:.2e} gold={s:.2e} mine={mn:.2e}" for e, k, s, mn in errs[:10]))
            assert errs[0][0] <= gtol, errs[0]
    finally:
        D.set_compute_dtype(torch.float32)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 3e-2)])
def test_drug_simclr_nt_xent_matches_oracle(dtype, tol):
    """SSL(drug_ssl_type='simclr').drug_simclr (model/self_supervised_learning.py:35-41,168-182: both
    projectors, then NT-Xent over the 2b x 2b similarity matrix without its diagonal) against the oracle:
    loss, input gradients and projector weight gradients."""
    import druglamp_b200 as D
    from druglamp_b200.ssl import SSL
    from oracle import restatement as R
    D.set_compute_dtype(dtype)
    try:
        B, Ld, C = 4, 6, 128
        g = torch.Generator().manual_seed(21)
        vd, xd = torch.randn(B, Ld, C, generator=g), torch.randn(B, Ld, C, generator=g)
        ssl = SSL(torch.nn.Identity(), 640, drug_ssl_type="simclr").cuda().train(True)
        with torch.no_grad():
            ssl.drug_simclr(vd.cuda(), xd.cuda())              # materialises the lazy projectors
        shapes = {k: tuple(v.shape) for k, v in ssl.state_dict().items()}
        sd = R.deterministic_state(shapes)
        ssl.load_state_dict(sd, strict=True)
        a, b = vd.cuda().requires_grad_(True), xd.cuda().requires_grad_(True)
        ssl.zero_grad()
        loss = ssl.drug_simclr(a, b)
        loss.backward()
        torch.cuda.synchronize()
        sdr = {"ssl_model." + k: v.clone() for k, v in sd.items()}
        for k, v in sdr.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
        ar, br = vd.clone().requires_grad_(True), xd.clone().requires_grad_(True)
        q = R._simsiam_mlp(sdr, "ssl_model.net.projector.", ar.reshape(-1, C), True)
        k_ = R._simsiam_mlp(sdr, "ssl_model.llm_net.projector.", br.reshape(-1, C), True)
        ref = R.nt_xent(q, k_, 0.1)
        ref.backward()
        assert abs(float(loss) - float(ref)) <= tol * abs(float(ref)), (float(loss), float(ref))

        def cos(x, y):
            x, y = x.detach().double().flatten().cpu(), y.detach().double().flatten()
            return float((x * y).sum() / (x.norm() * y.norm()).clamp_min(1e-30))
        # bf16: logits of magnitude ~100 (temperature 0.1) carry ~0.5 of rounding, the softmax is that sharp
        bar = 0.9999 if dtype == torch.float32 else 0.95
        assert cos(a.grad, ar.grad) >= bar and cos(b.grad, br.grad) >= bar
        for name in ("net.projector.0.weight", "net.projector.6.weight", "llm_net.projector.3.weight"):
            p = dict(ssl.named_parameters())[name]
            assert cos(p.grad, sdr["ssl_model." + name].grad) >= bar, name
    finally:
        D.set_compute_dtype(torch.float32)
