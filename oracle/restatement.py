"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the DrugLAMP hot path.

A plain-PyTorch fp32 *functional* restatement of every reference function on the
hot path, written from the reference's behaviour,
each function citing the reference file:line it follows.  It takes the reference's
own ``state_dict`` (name -> tensor), so identical weights feed the reference, this
oracle and the CUDA product.

Pinning: the reference ships no tests or golden vectors, so this
restatement is pinned against *outputs of the reference itself*, run in the build
container through ``oracle/ref_shim.py``: ``tests/golden/make_golden.py`` writes the
fixtures under ``tests/golden/`` and ``tests/test_oracle_vs_reference.py`` re-checks
live whenever ``/root/reference`` is present.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
reference arm may import this module.  The product never does.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- H1 / H2
def fill_bit_cat(x: torch.Tensor):
    """``DrugLAMP.py:11-19``: bit = float(x.sum(-1) == 0); x' = cat(x, bit)."""
    bit = (x.sum(dim=-1) == 0).to(x.dtype)
    return bit, torch.cat((x, bit.unsqueeze(-1)), dim=-1)


def site_pool(x: torch.Tensor, site_len: int = 9, n_sites: int = 256):
    """``DrugLAMP.py:35-40``: view(B, 9, 256, C).mean(1)."""
    return x.reshape(-1, site_len, n_sites, x.shape[-1]).mean(dim=1)


# --------------------------------------------------------------------------- BN helper
def _bn(sd: SD, p: str, x: torch.Tensor, training: bool, affine: bool = True):
    """nn.BatchNorm1d over dim 0 (or (N,C,L)); updates running stats like the module."""
    w = sd[p + "weight"] if affine else None
    b = sd[p + "bias"] if affine else None
    y = F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], w, b,
                     training, 0.1, 1e-5)
    if training and (p + "num_batches_tracked") in sd:
        sd[p + "num_batches_tracked"] += 1
    return y


# --------------------------------------------------------------------------- H3-H5
def molecular_gcn(sd: SD, p: str, src, dst, h, batch_size: int, training: bool):
    """``basic_model.py:147-153`` + ``GCNLayer.forward :411-436`` + ``GraphConv.forward
    :545-638`` (norm='both', aggregate-then-matmul branch).  DGL's
    update_all(copy_u,sum) is an unweighted sum over in-edges, duplicates counted."""
    n = h.shape[0]
    x = h @ sd[p + "init_transform.weight"].t()
    out_norm = torch.bincount(src, minlength=n).clamp(min=1).to(x.dtype).pow(-0.5)
    in_deg = torch.bincount(dst, minlength=n)
    if bool((in_deg == 0).any()):
        raise Exception("There are 0-in-degree nodes in the graph")
    in_norm = in_deg.clamp(min=1).to(x.dtype).pow(-0.5)
    n_layers = len([k for k in sd if k.startswith(p + "gnn.gnn_layers.") and k.endswith("graph_conv.weight")])
    for l in range(n_layers):
        q = f"{p}gnn.gnn_layers.{l}."
        feat_src = x * out_norm[:, None]
        agg = torch.zeros_like(x).index_add_(0, dst, feat_src[src])
        rst = agg @ sd[q + "graph_conv.weight"]            # weight is [in, out]
        rst = rst * in_norm[:, None] + sd[q + "graph_conv.bias"]
        rst = F.relu(rst)
        res = F.relu(F.linear(x, sd[q + "res_connection.weight"], sd[q + "res_connection.bias"]))
        x = _bn(sd, q + "bn_layer.", rst + res, training)
    return x.view(batch_size, -1, x.shape[-1])


# --------------------------------------------------------------------------- adjacent
def protein_cnn(sd: SD, p: str, v, fill_mask, training: bool):
    """``basic_model.py:172-180`` incl. the final ``.view`` reinterpretation (App. A4)."""
    e = F.embedding(v.long(), sd[p + "embedding.weight"], padding_idx=0)
    x = torch.cat((e, fill_mask.unsqueeze(-1)), dim=-1).transpose(2, 1)
    for i in (1, 2, 3):
        x = F.conv1d(x, sd[f"{p}conv{i}.weight"], sd[f"{p}conv{i}.bias"], padding="same")
        x = _bn(sd, f"{p}bn{i}.", F.relu(x), training)
    return x.reshape(x.size(0), x.size(2), -1)


def feed_forward_layer(sd: SD, p: str, x):
    """``basic_model.py:190-194``."""
    x = F.gelu(F.linear(x, sd[p + "lin1.weight"], sd[p + "lin1.bias"]))
    x = F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "lin2.weight"], sd[p + "lin2.bias"])


def mlp_head(sd: SD, p: str, x, training: bool):
    """``basic_model.py:210-215``."""
    for i in (1, 2, 3):
        x = _bn(sd, f"{p}bn{i}.", F.gelu(F.linear(x, sd[f"{p}fc{i}.weight"], sd[f"{p}fc{i}.bias"])), training)
    return F.linear(x, sd[p + "fc4.weight"], sd[p + "fc4.bias"])


def prot_adaptor(sd: SD, xp):
    """``DrugLAMP.py:43-48``."""
    xp = xp + feed_forward_layer(sd, "p_adaptor_wo_skip_connect.", xp)
    xp = F.gelu(F.linear(xp, sd["lin_p1.weight"], sd["lin_p1.bias"]))
    xp = F.layer_norm(xp, (xp.shape[-1],), sd["p_norm.weight"], sd["p_norm.bias"], 1e-5)
    return F.linear(xp, sd["lin_p2.weight"], sd["lin_p2.bias"])


def drug_adaptor(sd: SD, xd):
    """``DrugLAMP.py:50-52``."""
    xd = F.gelu(F.linear(xd, sd["lin_d1.weight"], sd["lin_d1.bias"]))
    xd = F.layer_norm(xd, (xd.shape[-1],), sd["d_norm.weight"], sd["d_norm.bias"], 1e-5)
    return F.linear(xd, sd["lin_d2.weight"], sd["lin_d2.bias"])


# --------------------------------------------------------------------------- H6
def pgca(sd: SD, p: str, query, key, value, num_heads: int = 1):
    """``PGCA/guided_cross_attention_model.py:124-329``.  All three in-proj branches
    (:138-190) compute q=W_q query+b_q, k=W_k key+b_k, v=W_v value+b_v; returns
    ``(attn_output (L,N,E), raw scaled logits (N,H,L,S))`` (:307,:316-320)."""
    L, N, E = query.shape
    S = key.shape[0]
    hd = E // num_heads
    W, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(query, W[:E], b[:E]) * (float(hd) ** -0.5)
    k = F.linear(key, W[E:2 * E], b[E:2 * E])
    v = F.linear(value, W[2 * E:], b[2 * E:])
    q = q.contiguous().view(L, N * num_heads, hd).transpose(0, 1)
    k = k.contiguous().view(S, N * num_heads, hd).transpose(0, 1)
    v = v.contiguous().view(S, N * num_heads, hd).transpose(0, 1)
    raw = torch.bmm(q, k.transpose(1, 2))
    o = torch.bmm(F.softmax(raw, dim=-1), v)
    o = o.transpose(0, 1).contiguous().view(L, N, E)
    o = F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])
    return o, raw.view(N, num_heads, L, S)


# --------------------------------------------------------------------------- H7
def mhla(sd: SD, p: str, v):
    """``PMMA/encoder.py:127-140``: softmax over the sequence axis, then gating through
    the memory-reinterpreting ``.view(B*H, L, head_dim)`` (App. A3)."""
    B, L, E = v.shape
    a = F.gelu(F.linear(v, sd[p + "lin1.weight"], sd[p + "lin1.bias"]))
    a = F.linear(a, sd[p + "lin2.weight"], sd[p + "lin2.bias"])     # (B, L, H)
    H = a.shape[-1]
    a = F.softmax(a, dim=1).transpose(1, 2).contiguous()            # (B, H, L)
    out = a.view(B * H, L, 1) * v.contiguous().view(B * H, L, E // H)
    return out.view(B, L, E)


# --------------------------------------------------------------------------- H8-H11
def _heads(x, H):
    B, L, D = x.shape
    return x.view(B, L, H, D // H).permute(0, 2, 1, 3)


def _merge(x):
    B, H, L, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, L, H * d)


def _sdpa(q, k, v):
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    return torch.matmul(F.softmax(s, dim=-1), v)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln6(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-6)


def _ffn(sd, p, x):
    """``PMMA/mlp.py:44-50`` with dropout off."""
    return _lin(sd, p + "fc2", F.gelu(_lin(sd, p + "fc1", x)))


def pmma(sd: SD, p: str, prot, mol, num_heads: int = 4):
    """``PMMA/paired_multi_model_attention_model.py:22-29`` -> ``embed.py:38-54`` ->
    ``encoder.py:41-56`` -> ``block.py:33-62`` -> ``attention.py:90-127``.
    Dropout is off (eval / p=0) -- parity is defined without it."""
    e = p + "embeddings."
    mol = _lin(sd, e + "mol_embeddings", mol) + sd[e + "pe_mol"]
    prot = prot + sd[e + "pe_prot"]                   # embed.py:50-51 (the Linear is dead)
    n_layers = len([k for k in sd if k.startswith(p + "encoder.layer_with_mol.") and k.endswith("attention_norm.weight")])
    x = None
    for i in range(n_layers):
        q = f"{p}encoder.layer_with_mol.{i}."
        a = q + "attn."
        if i < 2:
            hp, hm = prot, mol
            xp_, xm_ = _ln6(sd, q + "attention_norm", prot), _ln6(sd, q + "att_norm_mol", mol)
            qp, kp, vp_ = (_heads(_lin(sd, a + n, xp_), num_heads) for n in ("query", "key", "value"))
            qm, km, vm = (_heads(_lin(sd, a + n, xm_), num_heads) for n in ("query_mol", "key_mol", "value_mol"))
            ap = torch.cat((_merge(_sdpa(qp, kp, vp_)), _merge(_sdpa(qm, kp, vp_))), dim=-1)
            am = torch.cat((_merge(_sdpa(qm, km, vm)), _merge(_sdpa(qp, km, vm))), dim=-1)
            prot = _lin(sd, a + "out", _lin(sd, a + "fc", ap)) + hp
            mol = _lin(sd, a + "out_mol", _lin(sd, a + "fc_mol", am)) + hm
            prot = _ffn(sd, q + "ffn.", _ln6(sd, q + "ffn_norm", prot)) + prot
            mol = _ffn(sd, q + "ffn_mol.", _ln6(sd, q + "ffn_norm_mol", mol)) + mol
        else:
            if i == 2:
                x = torch.cat((prot, mol), dim=-1)
            y = _ln6(sd, q + "attention_norm", x)
            qq, kk, vv = (_heads(_lin(sd, a + n, y), num_heads) for n in ("query", "key", "value"))
            x = _lin(sd, a + "out", _merge(_sdpa(qq, kk, vv))) + x
            x = _ffn(sd, q + "ffn.", _ln6(sd, q + "ffn_norm", x)) + x
    if x is None:
        x = torch.cat((prot, mol), dim=-1)
    return _ln6(sd, p + "encoder.encoder_norm", x)


# --------------------------------------------------------------------------- H14
def binary_cross_entropy(pred_output, labels):
    """``basic_model.py:17-22``: sigmoid + BCELoss (log clamped at -100)."""
    n = torch.sigmoid(pred_output).squeeze(1)
    return n, F.binary_cross_entropy(n, labels.float())


# --------------------------------------------------------------------------- H12
def tanh_decay(m_ori: float, n_re: int, step: int) -> float:
    """``utils.py:559-560``."""
    return float(m_ori * (1 - np.tanh(2 * (1 - step / n_re))))


def cm_label_matrix(meta: List[dict]):
    """``cross_modality.py:138-150``: dedup by id, last occurrence wins, python dict order
    (first-seen order of the ids); G defaults to 0 (use_cm=True)."""
    pid2t = {m["Prot_ID"]: t for t, m in enumerate(meta)}
    did2t = {m["Drug_ID"]: t for t, m in enumerate(meta)}
    prow = {pid: i for i, pid in enumerate(pid2t)}
    dcol = {did: j for j, did in enumerate(did2t)}
    G = torch.zeros(len(prow), len(dcol), dtype=torch.int64)
    for m in meta:
        G[prow[m["Prot_ID"]], dcol[m["Drug_ID"]]] = int(m["Y"])
    return list(pid2t.values()), list(did2t.values()), G


def cm_triplet_dense(p_lat, d_lat, G, margin: float):
    """Dense form of ``ccpp_p_tri_loss`` (``cross_modality.py:15-47``) with
    ``d(x,y) = 1 - sigmoid(cos(x,y))`` (``utils.py:571-574``) and reduction='sum'/n_tri."""
    cos = F.cosine_similarity(p_lat[:, None, :], d_lat[None, :, :], dim=-1)
    S = torch.sigmoid(cos)
    pos = (G == 1)
    neg = (G == 0)
    total = p_lat.new_zeros(())
    n_tri = 0
    for i in range(G.shape[0]):
        pi, ni = pos[i].nonzero().flatten(), neg[i].nonzero().flatten()
        if len(pi) > 0 and len(ni) > 0:
            t = S[i, ni][None, :] - S[i, pi][:, None] + margin      # d_ap - d_an + m
            total = total + F.relu(t).sum()
            n_tri += len(pi) * len(ni)
        elif len(ni) > 0:
            s_self = torch.sigmoid(F.cosine_similarity(p_lat[i:i + 1], p_lat[i:i + 1], dim=-1))
            total = total + F.relu(S[i, ni] - s_self + margin).sum()
            n_tri += len(ni)
    return total / max(n_tri, 1)


def cross_modality(sd: SD, p: str, prot, aug_prot, drug, aug_drug, meta, margin: float, training: bool):
    """``cross_modality.py:129-164`` (+ Mean2Embed :166-171)."""
    pt, dt, G = cm_label_matrix(meta)

    def m2e(name, x):
        x = _bn(sd, f"{p}{name}.0.", x, training)
        return F.linear(F.relu(x), sd[f"{p}{name}.2.weight"], sd[f"{p}{name}.2.bias"])

    pe = torch.cat((m2e("prot2latent", prot[pt].mean(1)), m2e("aug_prot2latent", aug_prot[pt].mean(1))), -1)
    de = torch.cat((m2e("drug2latent", drug[dt].mean(1)), m2e("aug_drug2latent", aug_drug[dt].mean(1))), -1)
    pl = F.normalize(F.linear(pe, sd[p + "to_prot_latent.weight"]), dim=-1)
    dl = F.normalize(F.linear(de, sd[p + "to_drug_latent.weight"]), dim=-1)
    return cm_triplet_dense(pl, dl, G, margin)


# --------------------------------------------------------------------------- H13
def mlm_loss(sd: SD, seq, masked_seq, labels, xp_cat, fill_bit, mode: str, training: bool):
    """``self_supervised_learning.py:78-101`` given the sampled mask (labels/masked_seq)."""
    out = []
    if mode != "xp":
        emb = protein_cnn(sd, "protein_extractor.", masked_seq, fill_bit, training)
        logits = _lin(sd, "ssl_model.to_logits", emb)
        out.append(F.cross_entropy(logits.transpose(1, 2), labels, ignore_index=0))
    if mode != "vp":
        llm_logits = _lin(sd, "ssl_model.llm_to_logits", xp_cat)
        out.append(F.cross_entropy(llm_logits.transpose(1, 2), labels, ignore_index=0))
    return sum(out) / len(out)


def _simsiam_mlp(sd, p, x, training):
    """``self_supervised_learning.py:153-166``."""
    x = F.relu(_bn(sd, p + "1.", F.linear(x, sd[p + "0.weight"]), training))
    x = F.relu(_bn(sd, p + "4.", F.linear(x, sd[p + "3.weight"]), training))
    return _bn(sd, p + "7.", F.linear(x, sd[p + "6.weight"]), training, affine=False)


def _predictor(sd, p, x, training):
    """``self_supervised_learning.py:143-151``."""
    x = F.relu(_bn(sd, p + "1.", _lin(sd, p + "0", x), training))
    return _lin(sd, p + "3", x)


def drug_simsiam(sd: SD, vd, xd_cat, training: bool):
    """``self_supervised_learning.py:43-65``.  BN running stats are updated by four
    projector passes and two predictor passes, in this order, like the reference."""
    one, two = vd.reshape(-1, vd.shape[-1]), xd_cat.reshape(-1, xd_cat.shape[-1])
    p1 = _simsiam_mlp(sd, "ssl_model.net.projector.", one, training)
    p2 = _simsiam_mlp(sd, "ssl_model.llm_net.projector.", two, training)
    q1 = _predictor(sd, "ssl_model.predictor.", p1, training)
    q2 = _predictor(sd, "ssl_model.predictor.", p2, training)
    with torch.no_grad():
        t1 = _simsiam_mlp(sd, "ssl_model.net.projector.", one, training)
        t2 = _simsiam_mlp(sd, "ssl_model.llm_net.projector.", two, training)

    def lf(x, y):
        return 2 - 2 * (F.normalize(x, dim=-1) * F.normalize(y, dim=-1)).sum(-1)

    return (lf(q1, t2) + lf(q2, t1)).mean()


def nt_xent(queries, keys, temperature: float = 0.1):
    """``self_supervised_learning.py:168-182``."""
    b = queries.shape[0]
    n = 2 * b
    projs = torch.cat((queries, keys))
    logits = projs @ projs.t()
    logits = logits[~torch.eye(n, dtype=torch.bool)].reshape(n, n - 1) / temperature
    labels = torch.cat((torch.arange(b) + b - 1, torch.arange(b)))
    return F.cross_entropy(logits, labels, reduction="sum") / n


# --------------------------------------------------------------------------- full models
def druglamp_forward(sd: SD, kind: str, src, dst, h, batch_size, vp, xd, xp, training: bool,
                     site_len: int = 9, seq_len: int = 2304):
    """``DrugLAMP.py:8-79`` / ``DrugLAMP2C2P.py:8-90`` / ``DrugLAMPwoLLM.py:8-52``.
    Returns a dict of every intermediate at a hot-path boundary."""
    o = {}
    vd = molecular_gcn(sd, "drug_extractor.", src, dst, h, batch_size, training)
    bit_p, xp_cat = fill_bit_cat(xp)
    o.update(vd=vd, fill_bit_p=bit_p)
    n_sites = seq_len // site_len
    vpf = protein_cnn(sd, "protein_extractor.", vp, bit_p, training)
    vpf = site_pool(vpf, site_len, n_sites)
    o["vp"] = vpf
    if kind != "DrugLAMPwoLLM":
        bit_d, xd_cat = fill_bit_cat(xd)
        o.update(fill_bit_d=bit_d, xp_cat=xp_cat, xd_cat=xd_cat)
        xpp = prot_adaptor(sd, site_pool(xp_cat, site_len, n_sites))
        xdd = drug_adaptor(sd, xd_cat)
        o.update(xp=xpp, xd=xdd)
    mv, A_v = pgca(sd, "v_gca.", vpf.permute(1, 0, 2), vd.permute(1, 0, 2), vd.permute(1, 0, 2))
    mv = torch.cat((vpf, mv.permute(1, 0, 2)), 2)
    mv = F.layer_norm(mhla(sd, "v_mhla.", mv) + mv, (mv.shape[-1],), sd["v_gca_norm.weight"], sd["v_gca_norm.bias"], 1e-5)
    o.update(A_v_gca=A_v, mv=mv)
    if kind != "DrugLAMPwoLLM":
        mx, A_x = pgca(sd, "x_gca.", xpp.permute(1, 0, 2), xdd.permute(1, 0, 2), xdd.permute(1, 0, 2))
        mx = torch.cat((xpp, mx.permute(1, 0, 2)), 2)
        mx = F.layer_norm(mhla(sd, "x_mhla.", mx) + mx, (mx.shape[-1],), sd["x_gca_norm.weight"], sd["x_gca_norm.bias"], 1e-5)
        o.update(A_x_gca=A_x, mx=mx)
        f = pmma(sd, "pmma.", mx, mv)
    else:
        f = pmma(sd, "pmma.", mv, mv)
    o["f"] = f
    o["score"] = mlp_head(sd, "mlp_classifier.", f.mean(dim=1), training)
    return o


# --------------------------------------------------------------------------- deterministic params
def _hash_uniform(n: int, key: int) -> np.ndarray:
    """n reproducible pseudo-random float64 in [-1, 1): splitmix64 of (key, index) in wrapping
    uint64 arithmetic -- a pure function of its arguments on every machine."""
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(key & 0xFFFFFFFFFFFFFFFF)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / float(1 << 53)) - 1.0


def deterministic_state(shapes: Dict[str, tuple], seed: int = 0) -> SD:
    """Machine-independent parameter values: a pure arithmetic function of (name, index), so the
    build container, the GPU box and the fixtures agree without shipping 56 MB of weights.
    Matrices are full-rank pseudo-random (hashed uniform) with standard deviation 0.8 / sqrt(fan_in),
    i.e. a healthy, initialisation-like network.  (Round 1 used sinusoids of the flat index: every
    such matrix has rank <= 4, the activations collapsed onto a few directions and the
    normalisation layers amplified bf16 rounding to 20-70 % logit noise -- for PyTorch's own bf16
    autocast of the reference just as much as for the product -- which pinned nothing.)"""
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        n = int(np.prod(shp)) if len(shp) else 1
        # ssl_model.extractor.* are the same tensors as protein_extractor.* (App. B)
        hname = name.replace("ssl_model.extractor.", "protein_extractor.")
        hsh = (sum((i + 1) * ord(c) for i, c in enumerate(hname)) * 2654435761 + seed * 97) % 1000003
        vals = _hash_uniform(n, hsh * 0x100000001B3 + 0x5851F42D4C957F2D)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shp, dtype=torch.int64)
            continue
        if name.endswith("running_var"):
            vals = 1.0 + 0.25 * vals ** 2
        elif name.endswith("running_mean"):
            vals = 0.05 * vals
        elif len(shp) >= 2:
            fan_in = int(np.prod(shp[1:])) if "graph_conv.weight" not in name else shp[0]
            if ".pe_" in name:
                vals = 0.1 * vals
            else:
                vals = vals * (0.8 * math.sqrt(3.0) / math.sqrt(fan_in))
        elif name.endswith("weight"):      # norm gains
            vals = 1.0 + 0.1 * vals
        else:                               # biases
            vals = 0.05 * vals
        out[name] = torch.from_numpy(vals.astype(np.float32).reshape(shp))
    alias_state(out)
    return out


def alias_state(sd: SD) -> SD:
    """Make ``ssl_model.extractor.*`` the *same tensors* as ``protein_extractor.*``
    (the reference shares the module, ``basic_model.py:79-84``)."""
    for k in list(sd):
        if k.startswith("ssl_model.extractor."):
            sd[k] = sd["protein_extractor." + k[len("ssl_model.extractor."):]]
    return sd


# ---------------------------------------------------------------------------------------------
# Collate padding (utils.py:304-324), the checker for druglamp_b200.collate / dl_expand_rows
def tail_pad(blocks, maxsize: int):
    """utils.py:304-312: each sample's rows once, zeros after."""
    out = torch.zeros(len(blocks), maxsize, blocks[0].shape[-1])
    for i, a in enumerate(blocks):
        out[i, :a.shape[-2], :] = a
    return out


def repeat_pad(blocks, maxsize: int):
    """utils.py:314-324: each sample's rows tiled floor(maxsize / rows) times, zeros after."""
    out = torch.zeros(len(blocks), maxsize, blocks[0].shape[-1])
    for i, a in enumerate(blocks):
        n = a.shape[-2]
        for j in range(maxsize // n):
            out[i, j * n:(j + 1) * n, :] = a
    return out
