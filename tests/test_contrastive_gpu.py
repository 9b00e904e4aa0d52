"""GPU: the 2C2P step with global negatives (druglamp_b200/contrastive.py, BASELINE configs[2]).

The two-pass gradient-cache scheme (features -> contrastive loss on all rows -> backbone pass with
<pooled, d pooled>) must produce the gradient of ONE autograd graph that holds every micro-batch's
forward and the contrastive loss over their concatenated features; and the CUDA-graph replay must
equal the eager step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(dtype, n_micro=2, B=4):
    import druglamp_b200 as D
    from druglamp_b200.models import DrugLAMP2C2P
    from druglamp_b200.modules import CrossModality
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch
    D.set_compute_dtype(dtype)
    torch.manual_seed(7)
    m = DrugLAMP2C2P(384, 640).cuda()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train()
    raw = [make_batch(B, seed=40 + i, drugs_per_protein=2.0) for i in range(n_micro)]
    meta = [x for b in raw for x in b.meta]
    sbs = [StaticBatch(b, torch.device("cuda")) for b in raw]
    return m, sbs, CrossModality.prepare(meta)


def test_two_pass_gradient_equals_one_autograd_graph():
    import druglamp_b200 as D
    from druglamp_b200.contrastive import ContrastiveStep, _pool
    from druglamp_b200.modules import binary_cross_entropy
    try:
        m, sbs, targets = _setup(torch.float32)
        cs = ContrastiveStep(m, targets, n_local=8, cm_weight=0.7)
        total = cs.step(sbs, update=False)
        got = cs.flat.grad.clone()
        bn_after = m.protein_extractor.bn1.running_mean.clone()
        # one graph: every micro-batch forward (its own BatchNorm statistics), the contrastive loss over
        # the concatenated pooled features, summed with the mean classification loss
        for b_, _ in [(b, None) for b in m.buffers()]:
            pass
        cs.flat.zero_grad()
        feats, cls = [], 0.0
        for sb in sbs:
            out = m(*sb.model_inputs())
            cls = cls + binary_cross_entropy(out[4], sb.y)[1] / len(sbs)
            feats.append([_pool(out[3][k]) for k in ("prot", "aug_prot", "drug", "aug_drug")])
        allf = [torch.cat([f[i] for f in feats]) for i in range(4)]
        pl, dl = m.cm_model.latents_from_pooled(*allf, cs.targets)
        ref = cls + m.cm_model.loss_from_latents(pl, dl, cs.targets.G) * 0.7
        ref.backward()
        want = cs.flat.grad
        assert abs(float(total) - float(ref)) <= 1e-5 * abs(float(ref)), (float(total), float(ref))
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 2e-4 * scale, float((got - want).abs().max()) / scale
        assert float(got.abs().sum()) > 0
        del bn_after
    finally:
        D.set_compute_dtype(torch.float32)


def test_graph_replay_equals_eager_contrastive_step():
    import druglamp_b200 as D
    from druglamp_b200.contrastive import ContrastiveStep
    try:
        m, sbs, targets = _setup(torch.float32)
        cs = ContrastiveStep(m, targets, n_local=8)
        e = cs.step(sbs, update=False).clone()
        g_eager = cs.flat.grad.clone()
        cs.capture(sbs, n_micro=2)
        r = cs.step(sbs, graphs=True, update=False)
        scale = float(g_eager.abs().max())
        assert abs(float(e) - float(r)) <= 1e-5 * abs(float(e))
        assert float((cs.flat.grad - g_eager).abs().max()) <= 2e-4 * scale
        before = cs.flat.flat.clone()
        cs.step(sbs, graphs=True)                    # the optimiser graph moves the parameters
        assert float((cs.flat.flat - before).abs().max()) > 0
    finally:
        D.set_compute_dtype(torch.float32)
