"""Runs in its own process (patch_reference() rebinds names inside the reference's modules for the
whole interpreter): the reference's OWN model classes -- model/DrugLAMP*.py:forward, built by
DrugLAMPBase.__init__ -- on top of druglamp_b200.patch_reference(), on cuda:0, then the calls
trainer.py makes on them (trainer.py:196-213): model forward, binary_cross_entropy + backward,
ssl_model(**ssl_input), cm_model(**cm_input, meta).  Prints one JSON line of results for
tests/test_dropin_reference_gpu.py to compare with the fixtures of the unpatched reference.

    python tests/dropin_runner.py <kind> <fixture.npz> <f32|bf16>
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    kind, fixture, mode = sys.argv[1], sys.argv[2], sys.argv[3]
    import druglamp_b200 as D
    from druglamp_b200 import _lib
    from druglamp_b200.synth import make_batch
    from oracle import ref_shim, restatement as R
    from tests.util import load_golden, digest

    fx = load_golden(fixture)
    # the reference's own layers (ProteinCNN, adaptors, MLP head) run in torch here: full fp32, like the
    # CPU run that recorded the fixtures
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref_shim.install()
    D.patch_reference()
    D.set_compute_dtype(torch.float32 if mode == "f32" else torch.bfloat16)
    m = ref_shim.build_reference_model(kind)
    import model.basic_model as bm
    from druglamp_b200 import modules as M
    assert type(m.drug_extractor) is M.MolecularGCN and type(m.v_gca) is M.GuidedCrossAttention
    assert type(m.pmma) is M.PairedMultimodelAttention and type(m.cm_model) is M.CrossModality
    assert type(m.protein_extractor) is bm.ProteinCNN            # the reference's own, unpatched
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(R.deterministic_state(shapes), strict=True)
    m = m.cuda()
    training = bool(int(fx["meta_training"]))
    m.train(training)
    B, seed = int(fx["meta_B"]), int(fx["meta_seed"])
    b = make_batch(B, seed=seed)
    g = ref_shim.FakeGraph(b.graph.src.cuda(), b.graph.dst.cuda(), b.graph.num_nodes(), B,
                           b.graph.ndata["h"].clone().cuda())
    n0 = _lib.launch_count()
    vd, vp, ssl_input, cm_input, score = m(g, b.vp.cuda(), b.xd.cuda(), b.xp.cuda())   # trainer.py:196
    n, loss = bm.binary_cross_entropy(score, b.y.cuda())                               # trainer.py:199
    more = (cm_input is not None)
    loss.backward(retain_graph=True)                                                     # trainer.py:200 (SSL / CM follow)
    torch.cuda.synchronize()
    out = {"kind": kind, "mode": mode, "launches": _lib.launch_count() - n0,
           "score_dtype": str(score.dtype), "vd_dtype": str(vd.dtype),
           "score": score.detach().float().flatten().cpu().tolist(), "loss": float(loss),
           "mask_equal": bool(np.array_equal(ssl_input["fill_bit_p"].cpu().numpy().astype(np.uint8),
                                             fx["fill_bit_p"])),
           "A_v_gca_shape": list(m.A_v_gca.shape)}
    gerr = {}
    gmax = max(np.abs(fx[k][:64]).max() for k in fx if k.startswith("grad/"))
    params = dict(m.named_parameters())
    for k in fx:
        if k.startswith("grad/"):
            p = params[k[5:]]
            assert p.grad is not None, k
            a = digest(p.grad.float())
            kk = min(64, p.grad.numel())
            scale = max(np.abs(fx[k][:kk]).max(), fx[k][-1] / max(p.grad.numel(), 1), 1e-3 * gmax)
            gerr[k[5:]] = float(np.abs(a[:kk] - fx[k][:kk]).max() / scale)
    worst = max(gerr, key=gerr.get)
    out["worst_grad"] = [worst, gerr[worst]]
    # trainer.py:205 -- the SSL heads on the dict the forward handed back
    ssl_out = m.ssl_model(**ssl_input)
    out["ssl"] = [float(ssl_out["prot_ssl"]), float(ssl_out["drug_ssl"])]
    (0.1 * (ssl_out["prot_ssl"] + ssl_out["drug_ssl"])).backward(retain_graph=more)
    if cm_input is not None:
        # trainer.py:213 -- the 2C2P loss with the batch's meta dicts; then the margin schedule
        cm_loss = m.cm_model(**cm_input, meta=b.meta)
        out["cm_loss"] = float(cm_loss)
        cm_loss.backward()
        m.cm_model.step()
        out["cm_margin_after_step"] = float(m.cm_model.m_sch_loss_fn.margin)
    torch.cuda.synchronize()
    out["finite_grads"] = bool(all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None))
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
