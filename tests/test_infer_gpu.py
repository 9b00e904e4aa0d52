"""GPU: the graph-captured eval forward (druglamp_b200/infer.py, the validation/test scoring step of
trainer.py:256-292) replays to exactly the eager eval forward, matches the oracle's eval forward on the
same weights (probabilities within 1e-4 in fp32 mode, 2e-3 in bf16) and does not touch BatchNorm's running
statistics."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-3)])
def test_infer_graph_replay_matches_eager_and_oracle(dtype, tol):
    import druglamp_b200 as D
    from druglamp_b200.infer import InferStep
    from druglamp_b200.models import DrugLAMP
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch
    from oracle import restatement as R
    D.set_compute_dtype(dtype)
    try:
        torch.manual_seed(3)
        m = DrugLAMP(384, 640)
        g = torch.Generator().manual_seed(8)
        with torch.no_grad():                      # running statistics away from their (0, 1) init
            for k, v in m.state_dict().items():
                if k.endswith("running_mean"):
                    v.copy_(torch.randn(v.shape, generator=g) * 0.1)
                if k.endswith("running_var"):
                    v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
        m = m.cuda()
        st = InferStep(m)
        b = make_batch(6, seed=31)
        sb = StaticBatch(b, "cuda")
        buffers = {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "num_batches" in k}
        n_eager, loss_eager = st.eager(sb)
        n_eager, loss_eager = n_eager.clone(), loss_eager.clone()
        st.capture(sb)
        n, loss = st.replay(sb)
        assert st.launches_per_step > 50
        assert torch.equal(n, n_eager) and torch.equal(loss, loss_eager)
        for k, v in buffers.items():
            assert torch.equal(m.state_dict()[k], v), k

        # oracle eval forward on the CPU with the same state
        sd = R.alias_state({k: v.detach().cpu().clone() for k, v in m.state_dict().items()})
        with torch.no_grad():
            o = R.druglamp_forward(sd, "DrugLAMP", b.graph.src, b.graph.dst, b.graph.ndata["h"], len(b.y),
                                   b.vp, b.xd, b.xp, False)
            n_ref, loss_ref = R.binary_cross_entropy(o["score"], b.y)
        err = float((n.float().cpu() - n_ref.float()).abs().max())
        assert err <= tol, err
        assert abs(float(loss) - float(loss_ref)) <= tol * max(1.0, abs(float(loss_ref)))
    finally:
        D.set_compute_dtype(torch.bfloat16)
