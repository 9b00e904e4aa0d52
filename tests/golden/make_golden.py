"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs ``/root/reference``):

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so the pins are
outputs of the reference's own modules (imported through ``oracle/ref_shim.py``) on
seeded synthetic batches (``druglamp_b200.synth.make_batch``) with machine-independent
weights (``oracle.restatement.deterministic_state``).  Dropout probabilities are set to
0 (parity is defined without dropout).  Each fixture stores, per
tensor, 64 evenly spaced samples plus sum / abs-sum, which is enough to pin a result
without shipping megabytes.
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim, restatement as R  # noqa: E402
from druglamp_b200.synth import make_batch  # noqa: E402

N_SAMPLES = 64


def digest(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().flatten()
    n = t.numel()
    idx = torch.linspace(0, n - 1, min(N_SAMPLES, n), dtype=torch.float64).long().clamp_(max=n - 1)
    out = torch.zeros(N_SAMPLES + 2, dtype=torch.float64)
    out[: idx.numel()] = t[idx]
    out[-2] = t.sum()
    out[-1] = t.abs().sum()
    return out.numpy()


def build(kind):
    m = ref_shim.build_reference_model(kind)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def load_det(m):
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(R.deterministic_state(shapes), strict=True)
    return shapes


# Parameters whose COMPLETE reference gradients the batch-64 fixture keeps (every module family of
# the path: GCN, CNN, adaptors, PGCA, MHLA, PMMA paired + plain blocks, encoder norm, MLP head).
FULL_GRADS = [
    "drug_extractor.init_transform.weight", "drug_extractor.gnn.gnn_layers.0.graph_conv.weight",
    "drug_extractor.gnn.gnn_layers.2.res_connection.weight", "drug_extractor.gnn.gnn_layers.1.bn_layer.weight",
    "protein_extractor.conv2.bias", "protein_extractor.bn3.weight", "protein_extractor.embedding.weight",
    "lin_d2.weight", "p_norm.weight", "lin_p2.bias",
    "v_gca.in_proj_weight", "v_gca.out_proj.weight", "x_gca.in_proj_bias", "x_gca.out_proj.bias",
    "v_mhla.lin2.weight", "x_mhla.lin1.bias", "v_gca_norm.weight",
    "pmma.embeddings.pe_prot", "pmma.embeddings.mol_embeddings.bias",
    "pmma.encoder.layer_with_mol.0.attn.query.weight", "pmma.encoder.layer_with_mol.0.attn.key_mol.bias",
    "pmma.encoder.layer_with_mol.0.attn.fc_mol.bias", "pmma.encoder.layer_with_mol.1.attn.value.weight",
    "pmma.encoder.layer_with_mol.1.ffn.fc2.bias", "pmma.encoder.layer_with_mol.1.ffn_norm_mol.weight",
    "pmma.encoder.layer_with_mol.2.attn.query.bias", "pmma.encoder.layer_with_mol.2.attention_norm.weight",
    "pmma.encoder.layer_with_mol.3.attn.out.bias", "pmma.encoder.layer_with_mol.3.ffn.fc1.bias",
    "pmma.encoder.encoder_norm.weight", "mlp_classifier.fc3.bias", "mlp_classifier.fc4.weight",
    "mlp_classifier.bn1.weight",
]
FULL_PAIRS = (0, 63)     # pairs whose complete vd / vp rows are kept


def run_model(kind, B, seed, training, full=False):
    torch.manual_seed(0)
    m = build(kind)
    from model.basic_model import binary_cross_entropy
    shapes = load_det(m)
    m.train(training)
    b = make_batch(B, seed=seed)
    g = ref_shim.FakeGraph(b.graph.src, b.graph.dst, b.graph.num_nodes(), B, b.graph.ndata["h"].clone())
    vd, vp, ssl, cp, score = m(g, b.vp, b.xd, b.xp)
    n, loss = binary_cross_entropy(score, b.y)
    loss.backward()
    fx = {"score": score.detach().double().numpy(), "loss": np.float64(loss.item()),
          "prob": n.detach().double().numpy(),
          "vd": digest(vd), "vp": digest(vp), "A_v_gca": digest(m.A_v_gca)}
    if kind != "DrugLAMPwoLLM":
        fx["A_x_gca"] = digest(m.A_x_gca)
        fx["ssl_xp"] = digest(ssl["xp"])
        fx["ssl_xd"] = digest(ssl["xd"])
    fx["fill_bit_p"] = ssl["fill_bit_p"].detach().numpy().astype(np.uint8)
    if cp is not None:
        for k, v in cp.items():
            fx["cp_" + k] = digest(v)
    gnames = []
    for k, p in m.named_parameters():
        if p.grad is not None:
            fx["grad/" + k] = digest(p.grad)
            gnames.append(k)
    if full:
        # complete tensors, not digests: the bench configuration's parity pin (bf16 included)
        for k in FULL_GRADS:
            fx["fullgrad/" + k] = dict(m.named_parameters())[k].grad.detach().float().numpy()
        for i in FULL_PAIRS:
            fx[f"full_vd/{i}"] = vd[i].detach().float().numpy()
            fx[f"full_vp/{i}"] = vp[i].detach().float().numpy()
            fx[f"full_A_v_gca/{i}"] = m.A_v_gca[i].detach().float().numpy().astype(np.float16)
    for k, v in m.state_dict().items():
        if "running_" in k:
            fx["buf/" + k] = digest(v)
    # How far the UNMODIFIED reference moves when PyTorch itself runs it in bf16 (autocast): the
    # inherent bf16 sensitivity of this fixture (train-mode BatchNorm over a small batch amplifies
    # rounding noise).  The bf16 parity test bounds the product's deviation by this figure.
    m16 = build(kind)
    load_det(m16)
    m16.train(training)
    g16 = ref_shim.FakeGraph(b.graph.src, b.graph.dst, b.graph.num_nodes(), B, b.graph.ndata["h"].clone())
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        s16 = m16(g16, b.vp, b.xd, b.xp)[4].float()
    n16, l16 = binary_cross_entropy(s16, b.y)
    fx["bf16_autocast_score_dev"] = np.float64((s16 - score.detach()).abs().max() / score.detach().abs().max())
    fx["bf16_autocast_loss_dev"] = np.float64(abs(l16.item() - loss.item()) / abs(loss.item()))
    fx["meta_kind"] = np.array(kind)
    fx["meta_B"] = np.int64(B)
    fx["meta_seed"] = np.int64(seed)
    fx["meta_training"] = np.int64(int(training))
    return fx, m, b, cp, ssl


def run_cm(m, b, cp):
    fx = {}
    cm = copy.deepcopy(m.cm_model)
    cm.train(True)
    losses, margins = [], []
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in cp.items()}
    for step in range(4):
        for v in leaves.values():
            v.grad = None
        cm.zero_grad()
        margins.append(float(cm.m_sch_loss_fn.margin))
        l = cm(**leaves, meta=b.meta)
        l.backward()
        losses.append(float(l.item()))
        if step == 1:
            for k, v in leaves.items():
                fx["cm_grad_in/" + k] = digest(v.grad)
            for k, p in cm.named_parameters():
                fx["cm_grad/" + k] = digest(p.grad)
        cm.step()
    fx["cm_losses"] = np.array(losses)
    fx["cm_margins"] = np.array(margins)
    return fx


def run_ssl(m, ssl):
    """Materialise the lazy SimSiam projectors, reload deterministic weights, then record
    the sampled MLM mask (bit-exact target for the product's sampler on CPU) and losses."""
    ref_shim.install()
    from utils import mask_with_tokens, get_mask_subset_with_prob, prob_mask_like
    fx = {}
    sm = m.ssl_model
    inp = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in ssl.items()}
    torch.manual_seed(11)
    sm(**inp)                                   # creates projectors
    load_det(m)
    m.train(True)
    seq = inp["vp"]
    torch.manual_seed(12)
    no_mask = mask_with_tokens(seq, {0})
    mask = get_mask_subset_with_prob(~no_mask, 0.15)
    labels = seq.masked_fill(~mask, 0).long()
    rp = prob_mask_like(seq, 0.9)
    masked_seq = seq.clone().masked_fill(mask * rp, 26)
    torch.manual_seed(12)
    out = sm(**inp)
    fx["ssl_labels"] = labels.numpy().astype(np.int8)
    fx["ssl_masked_seq"] = masked_seq.numpy().astype(np.int8)
    fx["ssl_prot"] = np.float64(out["prot_ssl"].item())
    fx["ssl_drug"] = np.float64(float(out["drug_ssl"]))
    (out["prot_ssl"] + out["drug_ssl"]).backward()
    for k, p in sm.named_parameters():
        if p.grad is not None:
            fx["ssl_grad/" + k] = digest(p.grad)
    return fx


def main():
    torch.set_num_threads(os.cpu_count())
    which = set(sys.argv[1:]) or {"b64", "small"}
    if "b64" in which:
        # the configuration bench.py times: full DrugLAMP, 64 pairs, train mode (dropout 0 for parity)
        fx, *_ = run_model("DrugLAMP", 64, 21, True, full=True)
        np.savez_compressed(os.path.join(HERE, "druglamp_train_b64_full.npz"), **fx)
        print("druglamp train b64: loss", fx["loss"], "autocast dev", fx["bf16_autocast_score_dev"],
              fx["bf16_autocast_loss_dev"])
    if "small" not in which:
        return
    fx, m, b, cp, ssl = run_model("DrugLAMP2C2P", 16, 7, True)
    fx.update(run_cm(m, b, cp))
    np.savez_compressed(os.path.join(HERE, "druglamp2c2p_train_b16.npz"), **fx)
    print("2c2p train: loss", fx["loss"], "cm", fx["cm_losses"])

    fx, m, b, cp, ssl = run_model("DrugLAMP", 8, 9, True)
    fx.update(run_ssl(m, ssl))
    np.savez_compressed(os.path.join(HERE, "druglamp_train_b8_ssl.npz"), **fx)
    print("druglamp train+ssl: loss", fx["loss"], fx["ssl_prot"], fx["ssl_drug"])

    fx, *_ = run_model("DrugLAMP", 2, 3, False)
    np.savez_compressed(os.path.join(HERE, "druglamp_eval_b2.npz"), **fx)
    print("druglamp eval: loss", fx["loss"])

    fx, *_ = run_model("DrugLAMPwoLLM", 12, 5, True)
    np.savez_compressed(os.path.join(HERE, "druglampwollm_train_b12.npz"), **fx)
    print("wollm train: loss", fx["loss"])


if __name__ == "__main__":
    main()
