"""Host-side mirror of the reference's hot-path nn.Modules.

Same class names, constructor signatures, ``forward`` signatures, return tuples and ``state_dict``
keys/shapes as ``/root/reference/model``, so they drop into
``model/DrugLAMP*.py`` and ``trainer.py`` unchanged; the arithmetic is the hand-written sm_100a
kernels behind ``libdruglamp_sm100.so``.  There is no PyTorch fallback path: inputs must be CUDA
tensors and unsupported option combinations raise ``NotImplementedError`` instead of silently
running eager code.
"""
from __future__ import annotations

import copy
import math
import os
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functions as Fn
from . import kernels as K
from .graph import BatchedMolGraph

# druglamp_b200.patch_reference() turns this on: the drop-in modules then return their results in
# the CALLER's dtype.  Inside the reference's own model/DrugLAMP*.py the replaced modules alternate
# with the reference's fp32 nn.Linear / nn.LayerNorm / MLP layers, which reject a bf16 input
# ("mat1 and mat2 must have the same dtype"); druglamp_b200.models.* keeps everything in the compute
# dtype and leaves it off.
RETURN_CALLER_DTYPE = False


# Precision policy of the molecular GCN in bf16 mode.  92 % of its node rows are identical virtual
# nodes (App. A7), so every BatchNorm1d input has |mean| >> std: a bf16 rounding of the pre-norm
# activations is amplified by |mean| / std on the way through the three stacked train-mode norms
# (measured at 64 pairs: GCN weight-gradient cosines 0.96-0.98 against the reference).  The GCN is
# 1.3 % of the step's FLOPs, so by default its activations stay fp32 (TF32x3 GEMMs) whatever the
# compute dtype of the rest; "compute" follows set_compute_dtype.
GCN_PRECISION = "fp32"
# The decoder head (mean over the sequence -> MLP with three train-mode BatchNorm1d over the PAIRS of the
# batch, model/basic_model.py:196-215) works on 64 rows: 0.04 % of the FLOPs, but every BatchNorm over 64
# samples re-amplifies the rounding of its input.  It runs in fp32 as well.
HEAD_PRECISION = "fp32"
# Decoder head on <= 64 rows: one launch per layer (csrc/head.cu) instead of GEMM + BatchNorm chains.
HEAD_SMALL_KERNELS = os.environ.get("DL_NO_HEAD_KERNELS", "0") == "0"
# Evaluate the GCN once for all virtual nodes (graph.BatchedMolGraph.compact): ~12x fewer rows.
GCN_DEDUP = True


def set_gcn_precision(mode: str) -> None:
    global GCN_PRECISION
    if mode not in ("fp32", "compute"):
        raise ValueError("GCN precision must be 'fp32' or 'compute'")
    GCN_PRECISION = mode


def set_head_precision(mode: str) -> None:
    global HEAD_PRECISION
    if mode not in ("fp32", "compute"):
        raise ValueError("head precision must be 'fp32' or 'compute'")
    HEAD_PRECISION = mode


def _ret(y, like):
    if RETURN_CALLER_DTYPE and torch.is_tensor(y) and torch.is_tensor(like) and like.is_floating_point() \
            and y.dtype != like.dtype:
        return Fn.CastFn.apply(y, like.dtype)
    return y


# ================================================================================ PGCA (H6)
class GuidedCrossAttention(nn.Module):
    """Pocket-guided cross attention: nn.MultiheadAttention semantics that additionally returns the
    raw, scaled, pre-softmax logits ``(N, H, L, S)``
    (reference ``model/PGCA/guided_cross_attention_model.py:332-486``)."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False,
                 add_zero_attn=False, kdim=None, vdim=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.kdim = kdim if kdim is not None else embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self._qkv_same_embed_dim = self.kdim == embed_dim and self.vdim == embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        if not self._qkv_same_embed_dim or add_bias_kv or add_zero_attn or not bias:
            raise NotImplementedError("the sm_100a PGCA kernel path covers kdim=vdim=embed_dim, bias=True, "
                                      "no bias_kv / zero_attn (the only configuration DrugLAMP uses)")
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.register_parameter('q_proj_weight', None)
        self.register_parameter('k_proj_weight', None)
        self.register_parameter('v_proj_weight', None)
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        self.bias_k = self.bias_v = None
        self.add_zero_attn = add_zero_attn
        self._reset_parameters()

    def _reset_parameters(self):
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, need_raw=True,
                attn_mask=None):
        if key_padding_mask is not None or attn_mask is not None:
            raise NotImplementedError("masks are never passed on the DrugLAMP hot path (SURVEY section 0)")
        if self.training and self.dropout > 0:
            raise NotImplementedError("attention dropout is 0 on the DrugLAMP hot path")
        out, raw = Fn.PGCAFn.apply(query, key, value, self.in_proj_weight, self.in_proj_bias,
                                   self.out_proj.weight, self.out_proj.bias, self.num_heads)
        out = _ret(out, query)
        if need_weights:
            return out, _ret(raw, query)
        return out, None


# ================================================================================ MHLA (H7)
class MultiHeadLinearAttention(nn.Module):
    """reference ``model/PMMA/encoder.py:88-140``."""

    def __init__(self, d_model, nhead, d_diff=32, dropout=0.1, activation='tanh'):
        super().__init__()
        if activation != 'gelu':
            raise NotImplementedError("the sm_100a MHLA path implements activation='gelu' (DrugLAMP's setting)")
        self.act = nn.GELU()
        self.lin1 = nn.Linear(d_model, d_diff)
        self.lin2 = nn.Linear(d_diff, nhead)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)

    def _check(self):
        if self.training and (self.dropout1.p > 0 or self.dropout2.p > 0):
            raise NotImplementedError("MHLA dropout is 0 on the DrugLAMP hot path (mlha_dropout=0)")

    def forward(self, v):
        self._check()
        return _ret(Fn.MHLAFn.apply(v, self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias,
                                    None, None, 0.0), v)

    def forward_residual_norm(self, v, norm: nn.LayerNorm):
        """``norm(v + self(v))`` in one fused kernel pair (reference ``model/DrugLAMP.py:63-71``)."""
        self._check()
        return Fn.MHLAFn.apply(v, self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias,
                               norm.weight, norm.bias, norm.eps)


# ================================================================================ PMMA (H8-H11)
class Mlp(nn.Module):
    """reference ``model/PMMA/mlp.py:29-50``."""

    def __init__(self, config):
        super().__init__()
        self.fc1 = nn.Linear(config.hidden_size, config.hidden_size * 4)
        self.fc2 = nn.Linear(config.hidden_size * 4, config.hidden_size)
        self.act_fn = F.gelu
        self.dropout = nn.Dropout(config.transformer["dropout_rate"])
        nn.init.xavier_uniform_(self.fc1.weight)
        nn.init.xavier_uniform_(self.fc2.weight)
        nn.init.normal_(self.fc1.bias, std=1e-6)
        nn.init.normal_(self.fc2.bias, std=1e-6)

    def forward(self, x, residual=None):
        p = self.dropout.p if self.training else 0.0
        return Fn.ffn(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual, p)


class Attention(nn.Module):
    """reference ``model/PMMA/attention.py:9-127`` (plain MHSA and the paired mode)."""

    def __init__(self, config, vis, mm=True):
        super().__init__()
        self.vis = vis
        self.num_attention_heads = config.transformer["num_heads"]
        self.attn_head_size = int(config.hidden_size / self.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attn_head_size
        hs = config.hidden_size
        self.query = nn.Linear(hs, self.all_head_size)
        self.key = nn.Linear(hs, self.all_head_size)
        self.value = nn.Linear(hs, self.all_head_size)
        rate = config.transformer["attention_dropout_rate"]
        if mm:
            self.query_mol = nn.Linear(hs, self.all_head_size)
            self.key_mol = nn.Linear(hs, self.all_head_size)
            self.value_mol = nn.Linear(hs, self.all_head_size)
            self.out_mol = nn.Linear(hs, hs)
            self.attn_dropout_mol = nn.Dropout(rate)
            self.attn_dropout_pm = nn.Dropout(rate)
            self.attn_dropout_mp = nn.Dropout(rate)
            self.proj_dropout_mol = nn.Dropout(rate)
            self.fc = nn.Linear(hs * 2, hs)
            self.fc_mol = nn.Linear(hs * 2, hs)
        self.out = nn.Linear(hs, hs)
        self.attn_dropout = nn.Dropout(rate)
        self.proj_dropout = nn.Dropout(rate)
        self.softmax = nn.Softmax(dim=-1)
        if rate != 0:
            raise NotImplementedError("attention_dropout_rate is 0 on the DrugLAMP hot path")
        if vis:
            raise NotImplementedError("vis=True (returning probability maps) is not on the hot path "
                                      "(DrugLAMPBase builds PMMA with vis=False)")

    def forward(self, hidden_states, mol=None, residual=None, residual_mol=None, side=None):
        """side: the stream the molecule stream's tensors (`mol`, `residual_mol`) live on; the returned
        molecule output stays on it (PMMABlock / Encoder join)."""
        H = self.num_attention_heads
        scale = 1.0 / math.sqrt(self.attn_head_size)
        if mol is None:
            # query / key / value in one (B, L, 3E) buffer: one GEMM when the weights are adjacent
            # in the flat parameter store (fused_parameter_groups), three otherwise
            qkv = Fn.QKVProjFn.apply(hidden_states, self.query.weight, self.query.bias, self.key.weight,
                                     self.key.bias, self.value.weight, self.value.bias)
            o = Fn.SelfAttnCoreFn.apply(qkv, H, scale)
            attn = Fn.linear(o, self.out.weight, self.out.bias, residual=residual)
            return attn, None, None
        if hidden_states.shape[1] != mol.shape[1]:
            raise ValueError("paired attention needs equal sequence lengths (as in the reference)")
        # one (2, B, L, 3E) buffer: slab 0 = q/k/v of the protein stream, slab 1 = of the molecule
        # stream; set 0 = protein queries, set 1 = molecule queries; each stream's K/V is read once
        qkv = Fn.PairedQKVFn.apply(hidden_states, mol,
                                   self.query.weight, self.query.bias, self.key.weight, self.key.bias,
                                   self.value.weight, self.value.bias,
                                   self.query_mol.weight, self.query_mol.bias, self.key_mol.weight,
                                   self.key_mol.bias, self.value_mol.weight, self.value_mol.bias)
        # op = [A(q_prot), A(q_mol)] on the protein K/V = cat(attn, attn_p);
        # om = the same query sets on the molecule K/V = (attn_p, attn): swapped
        op, om = Fn.PairedAttnCoreFn.apply(qkv, H, scale)
        if side is None:
            tp = Fn.FcCatFn.apply(op, self.fc.weight, self.fc.bias, False)
            tm = Fn.FcCatFn.apply(om, self.fc_mol.weight, self.fc_mol.bias, True)
            attn_prot = Fn.linear(tp, self.out.weight, self.out.bias, residual=residual)
            attn_mol = Fn.linear(tm, self.out_mol.weight, self.out_mol.bias, residual=residual_mol)
            return attn_prot, attn_mol, None, None
        core_done = torch.cuda.current_stream().record_event()
        with torch.cuda.stream(side):
            side.wait_event(core_done)
            Fn.crosses(side, om)
            tm = Fn.FcCatFn.apply(om, self.fc_mol.weight, self.fc_mol.bias, True)
            attn_mol = Fn.linear(tm, self.out_mol.weight, self.out_mol.bias, residual=residual_mol)
        tp = Fn.FcCatFn.apply(op, self.fc.weight, self.fc.bias, False)
        attn_prot = Fn.linear(tp, self.out.weight, self.out.bias, residual=residual)
        return attn_prot, attn_mol, None, None

    def fused_parameter_groups(self):
        """Parameter lists a FlatParams store should lay out back to back (params.fused_group)."""
        g = [[self.query.weight, self.key.weight, self.value.weight],
             [self.query.bias, self.key.bias, self.value.bias]]
        if hasattr(self, "query_mol"):
            g += [[self.query_mol.weight, self.key_mol.weight, self.value_mol.weight],
                  [self.query_mol.bias, self.key_mol.bias, self.value_mol.bias]]
        return g


class PMMABlock(nn.Module):
    """reference ``model/PMMA/block.py:18-62``; residual adds are fused into GEMM epilogues."""

    def __init__(self, config, vis, mm=False):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.attention_norm = nn.LayerNorm(config.hidden_size, eps=1e-6)
        self.ffn_norm = nn.LayerNorm(config.hidden_size, eps=1e-6)
        if mm:
            self.att_norm_mol = nn.LayerNorm(config.hidden_size, eps=1e-6)
            self.ffn_norm_mol = nn.LayerNorm(config.hidden_size, eps=1e-6)
            self.ffn_mol = Mlp(config)
        self.ffn = Mlp(config)
        self.attn = Attention(config, vis, mm)

    @staticmethod
    def _ln(x, m):
        return Fn.layer_norm(x, m.weight, m.bias, m.eps)

    @staticmethod
    def _ln_res(x, m):
        """(LayerNorm(x), x): the skip connection leaves through the same Function, so its gradient
        is added inside the LayerNorm backward kernel."""
        return Fn.layer_norm_res(x, m.weight, m.bias, m.eps)

    def forward(self, prot, mol=None, side=None):
        """side: a second CUDA stream on which `mol` is ready; the molecule stream's LayerNorms, output
        projection and FFN are issued there, concurrent with the protein stream's on the caller's stream,
        and the returned `mol` stays on it (Encoder.forward joins after the last paired block)."""
        if mol is None:
            h, res = self._ln_res(prot, self.attention_norm)
            prot, w, gw = self.attn(h, residual=res)
            h, res = self._ln_res(prot, self.ffn_norm)
            prot = self.ffn(h, residual=res)
            return prot, w, gw
        if side is not None:
            with torch.cuda.stream(side):
                hm, rm = self._ln_res(mol, self.att_norm_mol)
            hp, rp = self._ln_res(prot, self.attention_norm)
            prot, mol, w, gw = self.attn(hp, hm, residual=rp, residual_mol=rm, side=side)
            with torch.cuda.stream(side):
                hm, rm = self._ln_res(mol, self.ffn_norm_mol)
                mol = self.ffn_mol(hm, residual=rm)
            hp, rp = self._ln_res(prot, self.ffn_norm)
            prot = self.ffn(hp, residual=rp)
            return prot, mol, w, gw
        hp, rp = self._ln_res(prot, self.attention_norm)
        hm, rm = self._ln_res(mol, self.att_norm_mol)
        prot, mol, w, gw = self.attn(hp, hm, residual=rp, residual_mol=rm)
        hp, rp = self._ln_res(prot, self.ffn_norm)
        prot = self.ffn(hp, residual=rp)
        hm, rm = self._ln_res(mol, self.ffn_norm_mol)
        mol = self.ffn_mol(hm, residual=rm)
        return prot, mol, w, gw


class Embeddings(nn.Module):
    """reference ``model/PMMA/embed.py:24-54``.  ``self.embedding`` exists for the state_dict but
    its output is discarded by the reference (``:50-51``), so the GEMM is not issued."""

    def __init__(self, config, mol_len):
        super().__init__()
        self.embedding = nn.Linear(config.hidden_size, config.hidden_size)
        self.mol_embeddings = nn.Linear(config.hidden_size, config.hidden_size)
        self.pe_prot = nn.Parameter(torch.zeros(1, config.feat_len, config.hidden_size))
        self.pe_mol = nn.Parameter(torch.zeros(1, mol_len, config.hidden_size))
        self.dropout_mol = nn.Dropout(config.transformer["dropout_rate"])
        self.dropout = nn.Dropout(config.transformer["dropout_rate"])

    def forward(self, prot, mol):
        p = self.dropout.p if self.training else 0.0
        mol_embeddings = None
        if mol is not None:
            t = Fn.linear(mol, self.mol_embeddings.weight, self.mol_embeddings.bias)
            mol_embeddings = Fn.AddPEFn.apply(t, self.pe_mol, p, Fn.next_seed() if p > 0 else 0)
        embeddings = Fn.AddPEFn.apply(prot, self.pe_prot, p, Fn.next_seed() if p > 0 else 0)
        return embeddings, mol_embeddings


class Encoder(nn.Module):
    """reference ``model/PMMA/encoder.py:26-56`` including the in-place doubling of
    ``config.hidden_size`` at layer 2 (``:37``)."""

    def __init__(self, config, vis):
        super().__init__()
        self.vis = vis
        self.layer_with_mol = nn.ModuleList()
        self.encoder_norm = nn.LayerNorm(config.hidden_size * 2, eps=1e-6)
        for i in range(config.transformer["num_p_plus_s_layers"]):
            if i < 2:
                layer = PMMABlock(config, vis, mm=True)
            else:
                if i == 2:
                    config.hidden_size = config.hidden_size * 2
                layer = PMMABlock(config, vis)
            self.layer_with_mol.append(copy.deepcopy(layer))

    def forward(self, hidden_states, mol=None):
        attn_weights: List = []
        guided_attn_weights: List = []
        # the molecule stream of the paired blocks runs on a second CUDA stream (Fn.branch_stream)
        side = Fn.branch_stream(hidden_states) if mol is not None else None
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            Fn.crosses(side, mol)
        for i, layer_block in enumerate(self.layer_with_mol):
            if i >= 2:
                if i == 2:
                    if side is not None:
                        torch.cuda.current_stream().wait_stream(side)
                        Fn.crosses(torch.cuda.current_stream(), mol)
                        side = None
                    hidden_states = Fn.cat_last(hidden_states, mol)
                hidden_states, weights, guided_weights = layer_block(hidden_states)
            else:
                hidden_states, mol, weights, guided_weights = layer_block(hidden_states, mol, side=side)
            if self.vis:
                attn_weights.append(weights)
                guided_attn_weights.append(guided_weights)
        if side is not None:                      # fewer than three layers: nothing joined the streams yet
            torch.cuda.current_stream().wait_stream(side)
        m = self.encoder_norm
        encoded = Fn.layer_norm(hidden_states, m.weight, m.bias, m.eps)
        return encoded, attn_weights, guided_attn_weights


class PairedMultimodelAttention(nn.Module):
    """reference ``model/PMMA/paired_multi_model_attention_model.py:15-29``."""

    def __init__(self, config, vis=True):
        super().__init__()
        self.embeddings = Embeddings(config, mol_len=config.mol_len)
        self.encoder = Encoder(config, vis)

    def forward(self, prot, mol=None):
        embedding_output, mol = self.embeddings(prot, mol)
        encoded, w, gw = self.encoder(embedding_output, mol)
        return _ret(encoded, prot), w, gw


# ================================================================================ GCN (H3-H5)
class GraphConv(nn.Module):
    """Parameter holder with the reference's layout: weight is ``[in, out]``
    (``model/basic_model.py:438-500``)."""

    def __init__(self, in_feats, out_feats, norm='both', weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        if norm != 'both' or not weight or not bias:
            raise NotImplementedError("the sm_100a GCN path implements norm='both' with weight and bias")
        if in_feats > out_feats:
            raise NotImplementedError("in_feats > out_feats (matmul-then-aggregate) is never taken by DrugLAMP")
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.empty(out_feats))
        self._activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        nn.init.zeros_(self.bias)

    def forward(self, graph, feat, weight=None):
        g = BatchedMolGraph.from_dgl(graph)
        if not self._allow_zero_in_degree:
            g.check_no_zero_in_degree()
        agg = Fn.SpmmFn.apply(feat, g)
        act = K.ACT_RELU if self._activation is not None else K.ACT_NONE
        return Fn.LinearFn.apply(agg, self.weight, self.bias, act, None, 0.0, 0, True, False)


class GCNLayer(nn.Module):
    """reference ``model/basic_model.py:342-436``: BN(GraphConv(g,h) + ReLU(Linear(h)))."""

    def __init__(self, in_feats, out_feats, gnn_norm='both', activation=None, residual=True,
                 batchnorm=True, dropout=0.0, allow_zero_in_degree=False):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("GCN dropout is 0 on the DrugLAMP hot path")
        self.activation = activation
        self.graph_conv = GraphConv(in_feats, out_feats, norm=gnn_norm, activation=activation,
                                    allow_zero_in_degree=allow_zero_in_degree)
        self.dropout = nn.Dropout(dropout)
        self.residual = residual
        if residual:
            self.res_connection = nn.Linear(in_feats, out_feats)
        self.bn = batchnorm
        if batchnorm:
            self.bn_layer = nn.BatchNorm1d(out_feats)

    def reset_parameters(self):
        self.graph_conv.reset_parameters()
        if self.residual:
            self.res_connection.reset_parameters()
        if self.bn:
            self.bn_layer.reset_parameters()

    def forward(self, g, feats, last_row_weight=1.0):
        new_feats = self.graph_conv(g, feats)
        if self.residual:
            act = K.ACT_RELU if self.activation is not None else K.ACT_NONE
            new_feats = Fn.linear(feats, self.res_connection.weight, self.res_connection.bias, act,
                                  residual=new_feats)
        if self.bn:
            new_feats = Fn.batch_norm(new_feats, self.bn_layer, last_row_weight=last_row_weight)
        return new_feats


class GCN(nn.Module):
    """reference ``model/basic_model.py:217-340``."""

    def __init__(self, in_feats, hidden_feats=None, gnn_norm=None, activation=None, residual=None,
                 batchnorm=None, dropout=None, allow_zero_in_degree=None):
        super().__init__()
        if hidden_feats is None:
            hidden_feats = [64, 64]
        n = len(hidden_feats)
        gnn_norm = gnn_norm or ['both'] * n
        activation = activation or [F.relu] * n
        residual = residual or [True] * n
        batchnorm = batchnorm or [True] * n
        dropout = dropout or [0.0] * n
        self.hidden_feats = hidden_feats
        self.gnn_layers = nn.ModuleList()
        for i in range(n):
            self.gnn_layers.append(GCNLayer(in_feats, hidden_feats[i], gnn_norm[i], activation[i],
                                            residual[i], batchnorm[i], dropout[i],
                                            allow_zero_in_degree=bool(allow_zero_in_degree)))
            in_feats = hidden_feats[i]

    def reset_parameters(self):
        for gnn in self.gnn_layers:
            gnn.reset_parameters()

    def forward(self, g, feats, last_row_weight=1.0):
        for gnn in self.gnn_layers:
            feats = gnn(g, feats, last_row_weight=last_row_weight)
        return feats


class MolecularGCN(nn.Module):
    """reference ``model/basic_model.py:137-153``.  Accepts a DGLGraph (duck-typed) or a
    :class:`BatchedMolGraph`; note the reference zeroes OUTPUT row 127 of ``init_transform``
    (``:141-143``), reproduced here."""

    def __init__(self, in_feats, dim_embedding=128, padding=True, hidden_feats=None, activation=None):
        super().__init__()
        self.init_transform = nn.Linear(in_feats, dim_embedding, bias=False)
        if padding:
            with torch.no_grad():
                self.init_transform.weight[-1].fill_(0)
        self.gnn = GCN(in_feats=dim_embedding, hidden_feats=hidden_feats, activation=activation)
        self.output_feats = hidden_feats[-1]

    def forward(self, batch_graph):
        node_feats = batch_graph.ndata.pop('h')
        g = BatchedMolGraph.from_dgl(batch_graph)
        h_in = node_feats
        fp32 = torch.float32 if GCN_PRECISION == "fp32" else None
        cg = g.compact() if GCN_DEDUP else None
        if cg is not None:
            # real nodes + ONE representative virtual node; so few rows that the fp32 island can
            # afford the 3xTF32 GEMMs
            with K.local_compute_dtype(fp32, precise=True):
                x = Fn.linear(node_feats.index_select(0, cg.gather), self.init_transform.weight)
                x = self.gnn(cg.graph, x, last_row_weight=float(cg.n_virtual))
            node_feats = Fn.ExpandVirtualFn.apply(x, cg.real_idx, cg.n_full)
        else:
            # dense fallback (a graph assembled on the device, or one without virtual nodes): the weight
            # gradients sum tens of thousands of identical rows, plain TF32 shows there -- 3xTF32 as well
            with K.local_compute_dtype(fp32, precise=True):
                node_feats = Fn.linear(node_feats, self.init_transform.weight)
                node_feats = self.gnn(g, node_feats)
        return _ret(node_feats, h_in).view(batch_graph.batch_size, -1, self.output_feats)


# ================================================================================ CrossModality (H12)
def tanh_decay(m_ori, n_re, step):
    """reference ``utils.py:559-560``."""
    return m_ori * (1 - math.tanh(2 * (1 - step / n_re)))


class MarginSchedule:
    """State machine of the reference's MarginScheduledLossFunction
    (``model/cross_modality.py:49-102``): margin = m_ori until the first step(), then
    ``tanh_decay``; reset when the step counter reaches n_re."""

    def __init__(self, m_ori=0.25, n_epoch=100, n_re=-1):
        self.m_ori = m_ori
        self.n_epoch = n_epoch
        self.n_re = int(n_epoch * 0.2) if n_re == -1 else n_re
        self._step = 0
        self.m_cur = m_ori

    @property
    def margin(self):
        return self.m_cur

    def step(self):
        self._step += 1
        if self._step == self.n_re:
            self.reset()
        else:
            self.m_cur = tanh_decay(self.m_ori, self.n_re, self._step)

    def reset(self):
        self._step = 0
        self.m_cur = tanh_decay(self.m_ori, self.n_re, self._step)


def Mean2Embed(hidden=128):
    """reference ``model/cross_modality.py:166-171`` (keys ``.0.`` BN and ``.2.`` Linear)."""
    return nn.Sequential(nn.BatchNorm1d(hidden), nn.ReLU(inplace=True), nn.Linear(hidden, hidden))


class CMTargets:
    """Host-side result of the reference's dict-of-dict label construction
    (``model/cross_modality.py:138-150``): unique-row indices (last occurrence wins, first-seen
    order) and the dense label matrix.  Build once per batch with ``CrossModality.prepare``."""

    def __init__(self, p_idx, d_idx, G):
        self.p_idx, self.d_idx, self.G = p_idx, d_idx, G

    def to(self, device, non_blocking=False):
        return CMTargets(self.p_idx.to(device, non_blocking=non_blocking),
                         self.d_idx.to(device, non_blocking=non_blocking),
                         self.G.to(device, non_blocking=non_blocking))


class CrossModality(nn.Module):
    """2C2P contrastive block (reference ``model/cross_modality.py:104-164``)."""

    def __init__(self, *, use_cm=True, hidden_size=128, max_margin=0.5, n_re=100, **kwargs):
        self.use_cm = use_cm
        super().__init__()
        if not use_cm:
            raise NotImplementedError("use_cm=False (default cell -1) is never used by DrugLAMP")
        self.prot2latent = Mean2Embed(hidden_size)
        self.aug_prot2latent = Mean2Embed(hidden_size)
        self.drug2latent = Mean2Embed(hidden_size)
        self.aug_drug2latent = Mean2Embed(hidden_size)
        self.to_prot_latent = nn.Linear(hidden_size * 2, hidden_size * 2, bias=False)
        self.to_drug_latent = nn.Linear(hidden_size * 2, hidden_size * 2, bias=False)
        self.m_sch_loss_fn = MarginSchedule(m_ori=max_margin, n_re=n_re)

    def step(self):
        self.m_sch_loss_fn.step()

    @staticmethod
    def prepare(meta) -> CMTargets:
        pid2t = {m['Prot_ID']: t for t, m in enumerate(meta)}
        did2t = {m['Drug_ID']: t for t, m in enumerate(meta)}
        prow = {pid: i for i, pid in enumerate(pid2t)}
        dcol = {did: j for j, did in enumerate(did2t)}
        G = torch.zeros((len(prow), len(dcol)), dtype=torch.int8)
        for m in meta:
            G[prow[m['Prot_ID']], dcol[m['Drug_ID']]] = int(m['Y'])
        return CMTargets(torch.tensor(list(pid2t.values()), dtype=torch.int64),
                         torch.tensor(list(did2t.values()), dtype=torch.int64), G)

    @staticmethod
    def _pool(seq):
        return Fn.seq_mean(seq)                                                           # mean over L

    @staticmethod
    def _embed(x, m2e):
        x = Fn.batch_norm(x, m2e[0])
        x = Fn.ActFn.apply(x, K.ACT_RELU)
        return Fn.linear(x, m2e[2].weight, m2e[2].bias)

    def latents_from_pooled(self, prot, aug_prot, drug, aug_drug, targets: CMTargets):
        """Unit-norm latents of the unique entities from the per-pair sequence means (B, hidden).
        Selecting the unique rows after the mean equals the reference's select-then-mean."""
        # index_select: its backward is an index_add_ (no sort, no host sync: CUDA-graph capturable)
        prot, aug_prot = prot.index_select(0, targets.p_idx), aug_prot.index_select(0, targets.p_idx)
        drug, aug_drug = drug.index_select(0, targets.d_idx), aug_drug.index_select(0, targets.d_idx)
        pe = torch.cat([self._embed(prot, self.prot2latent), self._embed(aug_prot, self.aug_prot2latent)], -1)
        de = torch.cat([self._embed(drug, self.drug2latent), self._embed(aug_drug, self.aug_drug2latent)], -1)
        pl = Fn.L2NormFn.apply(Fn.linear(pe, self.to_prot_latent.weight))
        dl = Fn.L2NormFn.apply(Fn.linear(de, self.to_drug_latent.weight))
        return pl, dl

    def latents(self, prot, aug_prot, drug, aug_drug, targets: CMTargets):
        """Unit-norm protein / drug latents of the unique in-batch entities."""
        return self.latents_from_pooled(self._pool(prot), self._pool(aug_prot), self._pool(drug),
                                        self._pool(aug_drug), targets)

    def loss_from_latents(self, pl, dl, G):
        cos = Fn.MatmulNTFn.apply(pl, dl)
        return Fn.CMTripletFn.apply(cos, G, float(self.m_sch_loss_fn.margin))

    def forward(self, prot, aug_prot, drug, aug_drug, meta):
        targets = meta if isinstance(meta, CMTargets) else self.prepare(meta)
        if targets.G.device != prot.device:
            targets = targets.to(prot.device)
        pl, dl = self.latents(prot, aug_prot, drug, aug_drug, targets)
        return self.loss_from_latents(pl, dl, targets.G)


# ================================================================================ losses (H14)
def binary_cross_entropy(pred_output, labels):
    """reference ``model/basic_model.py:17-22``: returns ``(sigmoid(score).squeeze(1), BCELoss)``."""
    return Fn.BCEFn.apply(pred_output, labels)


# ================================================================================ adjacent modules
# DL_NO_CNN_TAIL=1: BatchNorm apply, transpose and site mean as separate passes (A/B measurements)
CNN_TAIL_FUSED = os.environ.get("DL_NO_CNN_TAIL", "0") == "0"


class ProteinCNN(nn.Module):
    """reference ``model/basic_model.py:155-180`` -- adjacent to the hot path, on
    the dl_* kernels end to end: fused embedding gather + fill bit, implicit-GEMM convolutions,
    dl_batchnorm, and the final ``.view`` reinterpretation (App. A4)."""

    def __init__(self, embedding_dim, num_filters, kernel_size, padding=True):
        super().__init__()
        self.embedding = nn.Embedding(27, embedding_dim - 1, padding_idx=0 if padding else None)
        in_ch = [embedding_dim] + num_filters
        self.in_ch = in_ch[-1]
        for i in range(3):
            setattr(self, f"conv{i + 1}", nn.Conv1d(in_ch[i], in_ch[i + 1], kernel_size[i], padding='same'))
            setattr(self, f"bn{i + 1}", nn.BatchNorm1d(in_ch[i + 1]))

    def forward(self, v, fill_mask, site_len: int = 0):
        """Channels-last throughout: each Conv1d('same')+ReLU is one implicit-GEMM dl_gemm launch,
        each BatchNorm1d runs on the (B*L, C) rows with the dl_batchnorm kernels; the last one is
        applied while the result is laid out as the reference's (B, C, L) buffer, which is then
        reinterpreted like its ``.view(B, L, C)`` (App. A4).  site_len > 0 (not part of the reference
        signature): also take the site mean of model/DrugLAMP.py:35-37 and return (B, L / site_len, C)."""
        emb = self.embedding
        x = Fn.EmbedFillFn.apply(v, fill_mask, emb.weight, emb.padding_idx)    # (B, L, 128): gather + fill bit
        for i in (1, 2, 3):
            conv, bn = getattr(self, f"conv{i}"), getattr(self, f"bn{i}")
            # conv -> ReLU -> BN: the ReLU's backward mask rides on the BatchNorm backward kernel
            x = Fn.Conv1dSameFn.apply(x, conv.weight, conv.bias, True, False)
            if i == 3 and CNN_TAIL_FUSED and Fn.cnn_tail_ok(x, site_len):
                return Fn.cnn_tail(x, bn, relu_input=True, site_len=site_len)
            x = Fn.batch_norm(x, bn, relu_input=True)
        y = Fn.TransposeFn.apply(x)                                         # (B, C, L) like the reference
        y = y.view(y.size(0), y.size(2), -1)
        return Fn.SitePoolFn.apply(y, site_len) if site_len > 0 else y


class FeedForwardLayer(nn.Module):
    """reference ``model/basic_model.py:182-194`` (LLM adaptor; adjacent, but runs on dl_gemm)."""

    def __init__(self, d_in, d_h):
        super().__init__()
        self.lin1 = nn.Linear(d_in, d_h)
        self.lin2 = nn.Linear(d_h, d_in)
        self.act = nn.GELU()
        self.norm = nn.LayerNorm(d_h)

    def forward(self, x, residual=None):
        # an input that carries alignment padding (641 -> 648 zero-padded columns from
        # K.fillbit_pool) keeps it on the way out, so the next layer takes it without a copy
        keep = x.shape[-1] != self.lin1.in_features
        x = Fn.linear(x, self.lin1.weight, self.lin1.bias, K.ACT_GELU)
        x = Fn.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return Fn.linear(x, self.lin2.weight, self.lin2.bias, residual=residual, keep_pad=keep)


class MLP(nn.Module):
    """reference ``model/basic_model.py:196-215`` (decoder head; adjacent, runs on dl_gemm)."""

    def __init__(self, in_dim, hidden_dim, out_dim, binary=1):
        super().__init__()
        self.fc1 = nn.Linear(in_dim, hidden_dim)
        self.act1 = nn.GELU()
        self.bn1 = nn.BatchNorm1d(hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.act2 = nn.GELU()
        self.bn2 = nn.BatchNorm1d(hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, out_dim)
        self.act3 = nn.GELU()
        self.bn3 = nn.BatchNorm1d(out_dim)
        self.fc4 = nn.Linear(out_dim, binary)

    def forward(self, x):
        rows = x.numel() // x.shape[-1]
        if HEAD_PRECISION == "fp32" and x.is_cuda and rows <= K.SMALL_M and HEAD_SMALL_KERNELS:
            # <= 64 pairs: each fc -> GELU -> BatchNorm1d layer is one launch (fp32 CUDA-core FMA)
            for i in (1, 2, 3):
                x = Fn.head_layer(x, getattr(self, f"fc{i}"), K.ACT_GELU, getattr(self, f"bn{i}"))
            return Fn.head_layer(x, self.fc4)
        with K.local_compute_dtype(torch.float32 if HEAD_PRECISION == "fp32" else None):
            for i in (1, 2, 3):
                fc, bn = getattr(self, f"fc{i}"), getattr(self, f"bn{i}")
                x = Fn.batch_norm(Fn.linear(x, fc.weight, fc.bias, K.ACT_GELU), bn)
            return Fn.linear(x, self.fc4.weight, self.fc4.bias).float()
