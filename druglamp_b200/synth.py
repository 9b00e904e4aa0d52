"""Synthetic DTI batches laid out exactly as the reference collate delivers them.

Layouts follow ``utils.multimodality_collate_func`` (reference ``utils.py:326-334``),
``tail_pad``/``repeat_pad`` (``utils.py:304-324``), ``repeat_integer_label_protein``
(``utils.py:392-412``) and the virtual-node / double self-loop construction in
``handler/dataset.py:212-222``.  Distributions are the ones SURVEY.md section 8d fixes.
There is no network, RDKit, ESM or ChemBERTa here, so atom features and LLM
embeddings are seeded surrogates; padding rows are exact zeros like the reference's.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .graph import BatchedMolGraph

MAX_NODES = 512
SEQ_LEN = 9 * 256
NODE_FEATS = 75
MAX_RESIDUES = 1022


@dataclass
class Batch:
    graph: BatchedMolGraph
    vp: torch.Tensor          # (B, 2304) float64 tokens 0..25
    y: torch.Tensor           # (B,) int64 {0,1}
    xd: torch.Tensor          # (B, 512, n_drug_feature) f32
    xp: torch.Tensor          # (B, 2304, n_prot_feature) f32
    meta: List[dict]
    n_atoms: np.ndarray
    prot_len: np.ndarray

    def to(self, device, non_blocking=False):
        return Batch(self.graph.to(device), self.vp.to(device, non_blocking=non_blocking),
                     self.y.to(device, non_blocking=non_blocking),
                     self.xd.to(device, non_blocking=non_blocking),
                     self.xp.to(device, non_blocking=non_blocking), self.meta,
                     self.n_atoms, self.prot_len)

    def model_inputs(self):
        """(vd, vp, xd, xp) in the order ``Model.forward`` takes them (DrugLAMP.py:8)."""
        return self.graph, self.vp, self.xd, self.xp

    def llm_blocks(self):
        """The per-sample embedding rows BEFORE the collate pads them (``l['drug'].x`` /
        ``l['prot'].x`` in ``utils.multimodality_collate_func``): (drug blocks, protein blocks)."""
        d = [self.xd[b, :min(int(self.n_atoms[b]) + 2, self.xd.shape[1])] for b in range(self.xd.shape[0])]
        p = [self.xp[b, :int(self.prot_len[b]) + 2] for b in range(self.xp.shape[0])]
        return d, p


def _molecule_edges(rng: np.random.Generator, n: int):
    """Random spanning tree + round(0.1 n) ring closures, both directions."""
    src, dst = [], []
    for a in range(1, n):
        b = int(rng.integers(0, a))
        src += [a, b]
        dst += [b, a]
    for _ in range(int(round(0.1 * n))):
        a, b = (int(v) for v in rng.integers(0, n, size=2))
        if a != b:
            src += [a, b]
            dst += [b, a]
    return src, dst


def make_batch(batch_size: int, seed: int = 1234, n_drug_feature: int = 384,
               n_prot_feature: int = 640, n_unique_frac: float = 60 / 64,
               drugs_per_protein: Optional[float] = None,
               max_atoms: int = 290, min_prot: int = 50, max_prot: int = MAX_RESIDUES) -> Batch:
    rng = np.random.default_rng(seed)
    B = batch_size
    n_atoms = np.clip(np.round(rng.lognormal(np.log(24.0), 0.45, B)), 4, max_atoms).astype(np.int64)
    prot_len = np.clip(np.round(rng.lognormal(np.log(480.0), 0.6, B)), min_prot, max_prot).astype(np.int64)

    # identities: draw pair ids so that a batch has ~n_unique_frac unique proteins / drugs;
    # BindingDB-shaped batches use ~drugs_per_protein drugs per protein.
    if drugs_per_protein is None:
        n_p = max(1, int(round(B * n_unique_frac)))
        n_d = max(1, int(round(B * n_unique_frac)))
    else:
        n_p = max(1, int(round(B / drugs_per_protein)))
        n_d = max(1, int(round(B * n_unique_frac)))
    pid = rng.integers(0, n_p, B)
    did = rng.integers(0, n_d, B)
    # entity-level properties are shared by every occurrence of the same id
    n_atoms = n_atoms[did]      # n_d, n_p <= B
    prot_len = prot_len[pid]
    y = (rng.random(B) < 0.45).astype(np.int64)

    feats = np.zeros((B, MAX_NODES, NODE_FEATS), dtype=np.float32)
    src_all, dst_all = [], []
    for b in range(B):
        ent = np.random.default_rng([seed, 1, int(did[b])])
        n = int(n_atoms[b])
        cols = ent.integers(0, 74, size=(n, 8))
        feats[b, np.arange(n)[:, None], cols] = 1.0
        feats[b, n:, 74] = 1.0                      # virtual-node bit (dataset.py:219)
        s, d = _molecule_edges(ent, n)
        loops_real = list(range(n))                 # smiles_to_bigraph(add_self_loop=True)
        loops_all = list(range(MAX_NODES))          # v_d.add_self_loop() after padding
        s = np.asarray(s + loops_real + loops_all, dtype=np.int64) + b * MAX_NODES
        d = np.asarray(d + loops_real + loops_all, dtype=np.int64) + b * MAX_NODES
        src_all.append(s)
        dst_all.append(d)
    src = torch.from_numpy(np.concatenate(src_all))
    dst = torch.from_numpy(np.concatenate(dst_all))
    h = torch.from_numpy(feats.reshape(B * MAX_NODES, NODE_FEATS))
    graph = BatchedMolGraph(src, dst, B * MAX_NODES, B, h)

    vp = np.zeros((B, SEQ_LEN), dtype=np.float64)
    xp = torch.zeros(B, SEQ_LEN, n_prot_feature, dtype=torch.float32)
    xd = torch.zeros(B, MAX_NODES, n_drug_feature, dtype=torch.float32)
    for b in range(B):
        L = int(prot_len[b])
        ent = np.random.default_rng([seed, 2, int(pid[b])])
        toks = ent.integers(1, 26, L)
        gp = torch.Generator().manual_seed(int(seed) * 1000003 + 2 * int(pid[b]) + 1)
        emb = torch.randn(L + 2, n_prot_feature, generator=gp)
        for i in range(SEQ_LEN // (L + 2)):
            st = i * (L + 2)
            vp[b, st + 1: st + 1 + L] = toks       # utils.py:403-407
            xp[b, st: st + L + 2] = emb            # utils.py:319-323
        gd = torch.Generator().manual_seed(int(seed) * 1000003 + 2 * int(did[b]))
        rows = min(int(n_atoms[b]) + 2, MAX_NODES)
        xd[b, :rows] = torch.randn(rows, n_drug_feature, generator=gd)

    meta = [{"Drug_ID": f"D{int(did[b])}", "Prot_ID": f"P{int(pid[b])}", "Y": int(y[b])}
            for b in range(B)]
    # one label per (protein, drug) identity, as in a real dataset
    seen = {}
    for b in range(B):
        key = (meta[b]["Prot_ID"], meta[b]["Drug_ID"])
        if key in seen:
            y[b] = seen[key]
            meta[b]["Y"] = int(y[b])
        seen[key] = int(y[b])
    return Batch(graph, torch.from_numpy(vp), torch.from_numpy(y), xd, xp, meta, n_atoms, prot_len)


# reference utils.py:345-371 (CHARPROTSET): letter -> 1..25 in this order, everything else padding (0)
_PROT_ALPHABET = "ACBEDGFIHKMLONQPSRUTWVYXZ"


def batch_from_records(records, seed: int = 0, n_drug_feature: int = 384, n_prot_feature: int = 640) -> Batch:
    """A batch from dataset rows ``{"smiles", "protein", "y"}`` (the shipped CSV columns) without
    RDKit / DGL / LLM checkpoints: residue tokens and labels are the real ones (``utils.py:392-412``
    tiling, 1022-residue truncation of ``handler/dataset.py:36``); the heavy-atom count is read off
    the SMILES string; bond topology, atom features and the LLM embedding rows are synthetic but a
    pure function of the strings, so the same molecule / protein always gets the same tensors."""
    import zlib
    B = len(records)
    feats = np.zeros((B, MAX_NODES, NODE_FEATS), dtype=np.float32)
    src_all, dst_all = [], []
    vp = np.zeros((B, SEQ_LEN), dtype=np.float64)
    xp = torch.zeros(B, SEQ_LEN, n_prot_feature, dtype=torch.float32)
    xd = torch.zeros(B, MAX_NODES, n_drug_feature, dtype=torch.float32)
    n_atoms = np.zeros(B, dtype=np.int64)
    prot_len = np.zeros(B, dtype=np.int64)
    y = np.zeros(B, dtype=np.int64)
    meta = []
    for b, r in enumerate(records):
        smi, seq = r["smiles"], r["protein"][:MAX_RESIDUES]
        hs, hp = zlib.crc32(smi.encode()), zlib.crc32(seq.encode())
        # heavy atoms: element symbols start with an upper-case letter (H is implicit); Cl / Br count once
        n = sum(1 for i, c in enumerate(smi) if c.isupper() and c != "H") + sum(smi.count(a) for a in "cnos")
        n = int(np.clip(n, 4, 290))
        n_atoms[b] = n
        ent = np.random.default_rng([seed, 1, hs])
        cols = ent.integers(0, 74, size=(n, 8))
        feats[b, np.arange(n)[:, None], cols] = 1.0
        feats[b, n:, 74] = 1.0
        s, d = _molecule_edges(ent, n)
        loops_real, loops_all = list(range(n)), list(range(MAX_NODES))
        src_all.append(np.asarray(s + loops_real + loops_all, dtype=np.int64) + b * MAX_NODES)
        dst_all.append(np.asarray(d + loops_real + loops_all, dtype=np.int64) + b * MAX_NODES)
        L = len(seq)
        prot_len[b] = L
        toks = np.array([_PROT_ALPHABET.find(c.upper()) + 1 for c in seq], dtype=np.float64)
        gp = torch.Generator().manual_seed(int(seed) * 1000003 + 2 * hp + 1)
        emb = torch.randn(L + 2, n_prot_feature, generator=gp)
        for i in range(SEQ_LEN // (L + 2)):
            st = i * (L + 2)
            vp[b, st + 1: st + 1 + L] = toks
            xp[b, st: st + L + 2] = emb
        gd = torch.Generator().manual_seed(int(seed) * 1000003 + 2 * hs)
        rows = min(n + 2, MAX_NODES)
        xd[b, :rows] = torch.randn(rows, n_drug_feature, generator=gd)
        y[b] = int(r["y"])
        meta.append({"Drug_ID": f"D{hs}", "Prot_ID": f"P{hp}", "Y": int(r["y"])})
    graph = BatchedMolGraph(torch.from_numpy(np.concatenate(src_all)), torch.from_numpy(np.concatenate(dst_all)),
                            B * MAX_NODES, B, torch.from_numpy(feats.reshape(B * MAX_NODES, NODE_FEATS)))
    return Batch(graph, torch.from_numpy(vp), torch.from_numpy(y), xd, xp, meta, n_atoms, prot_len)
