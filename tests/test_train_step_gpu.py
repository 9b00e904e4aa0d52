"""GPU: the training step bench.py times (druglamp_b200/train.py).

* the CUDA-graph replay of a captured step is the same computation as the eager step;
* FlatAdamW over the flat parameter buffer follows torch.optim.AdamW (the optimiser the reference
  builds in main.py:158-160) applied to the same gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(dtype, seed=3):
    import druglamp_b200 as D
    from druglamp_b200.models import DrugLAMP
    D.set_compute_dtype(dtype)
    torch.manual_seed(seed)
    m = DrugLAMP(384, 640).cuda()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0                      # replayed graphs bake the dropout seed in
    m.train()
    return m


def test_graph_replay_equals_eager_steps_and_flat_adamw_follows_torch_adamw():
    import druglamp_b200 as D
    from druglamp_b200.modules import binary_cross_entropy
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch, TrainStep
    try:
        b = make_batch(8, seed=5)
        sb = StaticBatch(b, torch.device("cuda"))
        # (a) five eager steps
        ma = _model(torch.float32)
        ta = TrainStep(ma, lr=1e-3, weight_decay=1e-2)
        la = [float(ta.eager(sb)) for _ in range(5)]
        # (b) capture + five replays.  Capturing (two eager warm-up steps on this very batch) must
        # leave parameters, Adam moments, step counter and BatchNorm buffers where they were.
        mb = _model(torch.float32)
        tb = TrainStep(mb, lr=1e-3, weight_decay=1e-2)
        before = tb.flat.flat.clone()
        bn_before = mb.mlp_classifier.bn1.running_mean.clone()
        tb.capture(sb, warmup=2)
        assert torch.equal(tb.flat.flat, before) and int(tb.opt.step_count) == 0
        assert torch.equal(mb.mlp_classifier.bn1.running_mean, bn_before)
        assert float(tb.opt.exp_avg.abs().sum()) == 0.0
        lb = [float(tb.replay(sb)) for _ in range(5)]
        # (lr 1e-3 on 8 pairs is a stiff problem: the losses swing 0.73 -> 0.02 in five steps and
        # summation-order differences of the split-K atomics grow to ~4e-4 by the last steps)
        assert la[0] > 0 and all(abs(x - y) <= 5e-3 * abs(x) for x, y in zip(la, lb)), (la, lb)
        # Adam's update is sign-like: where a gradient is ~0 its rounding noise decides the direction
        # of an lr-sized move, so single parameters may differ by a few lr; the bulk must agree
        d = (ta.flat.flat - tb.flat.flat).abs()
        assert float(d.mean()) <= 2e-5 and float((d > 1e-3).float().mean()) <= 2e-3, (float(d.mean()), float(d.max()))
        # (c) plain autograd + torch.optim.AdamW on an identically initialised, un-flattened model
        mc = _model(torch.float32)
        opt = torch.optim.AdamW(mc.parameters(), lr=1e-3, weight_decay=1e-2)
        lc = []
        for _ in range(5):
            opt.zero_grad(set_to_none=True)
            out = mc(*sb.model_inputs())
            _, loss = binary_cross_entropy(out[4], sb.y)
            loss.backward()
            opt.step()
            lc.append(float(loss.detach()))
        assert all(abs(x - y) <= 1e-2 * abs(x) for x, y in zip(la, lc)), (la, lc)
        pa = dict(ma.named_parameters())
        tot, big, n = 0.0, 0.0, 0
        skipped = 0
        for name, p in mc.named_parameters():
            if p.grad is None:
                # no gradient from the classification loss (SSL / CM heads, the dead PMMA embedding):
                # torch.optim.AdamW leaves them alone -- no decay either -- and so must FlatAdamW
                assert torch.equal(pa[name].detach(), p.detach()), name
                skipped += p.numel()
                continue
            dd = (pa[name] - p).abs()
            tot += float(dd.sum()); big += float((dd > 1e-3).sum()); n += dd.numel()
        assert skipped > 100_000
        assert tot / n <= 5e-5 and big / n <= 5e-3, (tot / n, big / n)
    finally:
        D.set_compute_dtype(torch.float32)
