"""CPU: the oracle restatement reproduces the fixtures generated from the unmodified
reference (tests/golden/make_golden.py), forward, loss, gradients and BN buffers."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from druglamp_b200.synth import make_batch
from tests.util import load_golden, assert_digest_close

CASES = ["druglamp2c2p_train_b16.npz", "druglamp_eval_b2.npz", "druglampwollm_train_b12.npz"]


def _state_for(fx, extra=()):
    names = [k[len("grad/"):] for k in fx if k.startswith("grad/")]
    return names


def oracle_run(fx, shapes):
    kind = str(fx["meta_kind"]); B = int(fx["meta_B"]); seed = int(fx["meta_seed"])
    training = bool(int(fx["meta_training"]))
    sd = R.deterministic_state(shapes)
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    b = make_batch(B, seed=seed)
    o = R.druglamp_forward(sd, kind, b.graph.src, b.graph.dst, b.graph.ndata["h"], B,
                           b.vp, b.xd, b.xp, training)
    n, loss = R.binary_cross_entropy(o["score"], b.y)
    loss.backward()
    return sd, b, o, n, loss


@pytest.mark.parametrize("case", CASES)
def test_restatement_matches_reference_fixture(case, model_shapes):
    fx = load_golden(case)
    sd, b, o, n, loss = oracle_run(fx, model_shapes)
    assert np.allclose(o["score"].detach().numpy(), fx["score"], rtol=1e-3, atol=2e-4)
    assert abs(loss.item() - float(fx["loss"])) < 5e-4 * max(1.0, abs(float(fx["loss"])))
    assert np.array_equal(o["fill_bit_p"].numpy().astype(np.uint8), fx["fill_bit_p"])     # bit-exact mask
    for k in ("vd", "vp", "A_v_gca"):
        assert_digest_close(o[k], fx[k], 1e-4, k)
    if "A_x_gca" in fx:
        assert_digest_close(o["A_x_gca"], fx["A_x_gca"], 1e-4, "A_x_gca")
        assert_digest_close(o["xp_cat"], fx["ssl_xp"], 1e-6, "xp_cat")
        assert_digest_close(o["xd_cat"], fx["ssl_xd"], 1e-6, "xd_cat")
    gmax = max(np.abs(fx[k][:64]).max() for k in fx if k.startswith("grad/"))
    for k in fx:
        if k.startswith("grad/"):
            g = sd[k[5:]].grad
            assert g is not None, k
            assert_digest_close(g, fx[k], 5e-3, k, floor=1e-4 * gmax)
        if k.startswith("buf/"):
            assert_digest_close(sd[k[4:]], fx[k], 1e-4, k)


def test_cm_losses_and_margin_schedule(model_shapes):
    fx = load_golden("druglamp2c2p_train_b16.npz")
    sd, b, o, n, loss = oracle_run(fx, model_shapes)
    margins = [0.5] + [R.tanh_decay(0.5, 100, s) for s in (1, 2, 3)]
    assert np.allclose(margins, fx["cm_margins"], rtol=1e-12)
    for step, m in enumerate(margins):
        sd2 = R.deterministic_state(model_shapes)
        for name in ("prot2latent", "aug_prot2latent", "drug2latent", "aug_drug2latent"):
            sd2[f"cm_model.{name}.0.num_batches_tracked"] += step
        # BN running stats do not influence the training-mode loss
        l = R.cross_modality(sd2, "cm_model.", o["vp"].detach(), o["xp"].detach(), o["vd"].detach(),
                             o["xd"].detach(), b.meta, m, True)
        assert abs(l.item() - fx["cm_losses"][step]) < 2e-4 * max(abs(fx["cm_losses"][step]), 1e-3), step


def _nt_xent_inputs():
    """Closed-form (RNG-free) inputs of the NT-Xent known-answer test."""
    b, d = 6, 16
    i = torch.arange(b * d, dtype=torch.float64)
    return torch.sin(0.37 * i + 0.1).view(b, d).float(), torch.cos(0.23 * i - 0.4).view(b, d).float()


def test_nt_xent_known_answer_from_the_reference():
    """model/self_supervised_learning.py:168-182 (`nt_xent_loss`) evaluated by the unmodified reference on
    the closed-form inputs above returned 61.0732421875 (generated in the build container through
    oracle/ref_shim.py; tests/test_oracle_vs_reference.py re-checks it live)."""
    q, k = _nt_xent_inputs()
    assert abs(float(R.nt_xent(q, k, 0.1)) - 61.0732421875) <= 1e-4
