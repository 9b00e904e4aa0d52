"""GPU: druglamp_b200.trainer_step.TrainerStep against the reference's own training step.

The reference side is ``trainer.py:179-231`` written out literally over the UNMODIFIED reference model
(``model/DrugLAMP2C2P.py`` from /root/reference or the byte-compiled oracle/_ref) on the CPU with three
``torch.optim.AdamW`` over all parameters (``main.py:158-160``): classification-only, SSL and SSL + 2C2P
steps in sequence.  What has to agree after every step is the parameter trajectory: WHICH parameters
moved at all (the gradient every optimiser sees is the last loss's, parameters it does not reach are
skipped) and the update itself (Adam normalises the gradient, so near-zero components may flip sign under
any re-rounding: per-tensor cosine of the update, not an element-wise bound)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

LRS = (1e-4, 5e-5, 2e-4)          # three different rates so the three optimisers are distinguishable


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def test_trainer_step_follows_the_reference_training_step():
    from oracle import ref_shim, restatement as R
    if not ref_shim.available():
        pytest.skip("neither /root/reference nor oracle/_ref is present")
    import druglamp_b200 as D
    from druglamp_b200 import models
    from druglamp_b200.ssl import sample_mlm_mask
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch
    from druglamp_b200.trainer_step import TrainerStep

    B = 6
    ref = ref_shim.build_reference_model("DrugLAMP2C2P")
    from model.basic_model import binary_cross_entropy as ref_bce
    D.set_compute_dtype(torch.float32)
    try:
        mine = models.DrugLAMP2C2P(384, 640).cuda()
        for m in (ref, mine):
            for mod in m.modules():
                if isinstance(mod, torch.nn.Dropout):
                    mod.p = 0.0
            m.train()
        # optimisers / flat store exist BEFORE the first SSL call creates the SimSiam projectors (main.py:158-160)
        opts = [torch.optim.AdamW(ref.parameters(), lr=lr) for lr in LRS]
        ts = TrainerStep(mine, *LRS)
        # materialise the lazy projectors on both sides, then give both the same deterministic weights
        b0 = make_batch(B, seed=70)
        with torch.no_grad():
            g0 = ref_shim.FakeGraph(b0.graph.src, b0.graph.dst, b0.graph.num_nodes(), B, b0.graph.ndata["h"].clone())
            ref.ssl_model(**ref(g0, b0.vp, b0.xd, b0.xp)[2])
            mine.ssl_model(**mine(*b0.to("cuda").model_inputs())[2])
        shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        assert shapes == {k: tuple(v.shape) for k, v in mine.state_dict().items()}
        sd = R.deterministic_state(shapes)
        ref.load_state_dict(sd, strict=True)
        mine.load_state_dict(sd, strict=True)
        ts.flat.sync(force=True)

        schedule = [(False, False), (True, False), (True, True), (False, False), (True, True)]
        names = [n for n, _ in ref.named_parameters()]
        for it, (ssl, cm) in enumerate(schedule):
            b = make_batch(B, seed=71 + it)
            before = {n: p.detach().clone() for n, p in ref.named_parameters()}
            mine_before = {n: p.detach().clone() for n, p in mine.named_parameters()}
            # ---- reference: trainer.py:195-229 ------------------------------------------------------
            g = ref_shim.FakeGraph(b.graph.src, b.graph.dst, b.graph.num_nodes(), B, b.graph.ndata["h"].clone())
            _, _, ssl_input, cm_input, score = ref(g, b.vp, b.xd, b.xp)
            opts[0].zero_grad()
            _, cls_loss = ref_bce(score, b.y)
            cls_loss.backward(retain_graph=ssl or cm)
            if ssl:
                opts[1].zero_grad()
                torch.manual_seed(500 + it)
                d = ref.ssl_model(**ssl_input)
                ssl_loss = (d["prot_ssl"] + d["drug_ssl"]) * 0.1
                ssl_loss.backward(retain_graph=cm)
            if cm:
                opts[2].zero_grad()
                cm_input["meta"] = b.meta
                cm_loss = ref.cm_model(**cm_input) * 1.0
                cm_loss.backward()
            opts[0].step()
            if ssl:
                opts[1].step()
            if cm:
                opts[2].step()
            # ---- product ------------------------------------------------------------------------------
            torch.manual_seed(500 + it)
            mask = sample_mlm_mask(b.vp) if ssl else None
            losses = ts.step(StaticBatch(b, "cuda"), meta=b.meta, compute_ssl=ssl, compute_cm=cm, mlm_mask=mask)
            torch.cuda.synchronize()
            # step 0 starts from identical weights; afterwards the two trajectories differ by what Adam makes
            # of rounding noise (sign-normalised updates of zero-gradient components), so the bar widens
            ltol = 1e-3 if it == 0 else 2e-2
            assert abs(float(losses["train_loss"]) - float(cls_loss)) <= ltol * abs(float(cls_loss)), it
            if ssl:
                assert abs(float(losses["ssl_loss"]) - float(ssl_loss)) <= max(ltol, 2e-3) * abs(float(ssl_loss)), it
            if cm:
                assert abs(float(losses["cm_loss"]) - float(cm_loss)) <= max(ltol, 2e-3) * abs(float(cm_loss)) + 1e-6, it
            # ---- the trajectory -----------------------------------------------------------------------
            mp = dict(mine.named_parameters())
            rp = dict(ref.named_parameters())
            gmax = max(float(p.grad.abs().max()) for p in rp.values() if p.grad is not None)
            mod_gmax = {}
            for n, p in rp.items():
                if p.grad is not None:
                    k = n.rsplit(".", 1)[0]
                    mod_gmax[k] = max(mod_gmax.get(k, 0.0), float(p.grad.abs().max()))
            moved_ref, moved_mine, low = set(), set(), []
            num = den_a = den_b = 0.0
            for n in names:
                dr = (rp[n].detach() - before[n])
                dm = (mp[n].detach() - mine_before[n]).cpu()
                if float(dr.abs().max()) > 0:
                    moved_ref.add(n)
                if float(dm.abs().max()) > 0:
                    moved_mine.add(n)
                # a gradient that is zero in exact arithmetic (key biases: softmax is shift invariant; MHLA's
                # lin2 bias likewise; a bias in front of a train-mode BatchNorm) is rounding noise that Adam
                # normalises to +-lr: no direction to compare.  Recognised by its size next to the gradient of
                # the weight it belongs to.
                noise = rp[n].grad is not None and (
                    float(rp[n].grad.abs().max()) < 1e-5 * gmax or
                    float(rp[n].grad.abs().max()) < 1e-3 * mod_gmax[n.rsplit(".", 1)[0]])
                if n in moved_ref and n in moved_mine and not noise:
                    c = _cos(dm, dr)
                    num += float(dm.double().flatten() @ dr.double().flatten())
                    den_a += float(dm.double().pow(2).sum()); den_b += float(dr.double().pow(2).sum())
                    if c < 0.9 and "protein_extractor" not in n:      # (ReLU-mask flips, see test_model_parity_gpu.py)
                        low.append((n, round(c, 4)))
            assert moved_ref == moved_mine, (it, sorted(moved_ref ^ moved_mine)[:10])
            assert num / (den_a * den_b) ** 0.5 >= 0.99, (it, num / (den_a * den_b) ** 0.5)
            assert not low, (it, low[:10])
            if cm:       # the 2C2P gradient is the one applied: nothing behind the 2C2P inputs moves
                assert not any(n.startswith(("pmma.", "mlp_classifier.", "v_gca.", "ssl_model.to_logits")) for n in moved_mine)
            elif ssl:
                assert not any(n.startswith(("pmma.", "mlp_classifier.", "cm_model.")) for n in moved_mine)
            else:
                assert any(n.startswith("pmma.") for n in moved_mine) and not any(n.startswith("cm_model.") for n in moved_mine)
    finally:
        D.set_compute_dtype(torch.float32)
