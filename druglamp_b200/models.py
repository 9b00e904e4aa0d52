"""DrugLAMP model variants on the sm_100a kernels -- same constructors, ``forward`` signatures,
return tuples, attributes and ``state_dict`` as reference ``model/DrugLAMP.py``,
``model/DrugLAMPwoLLM.py``, ``model/DrugLAMP2C2P.py`` and ``model/basic_model.py:57-135``.

Differences from running the reference classes over the drop-in modules (which also works, see
``druglamp_b200.patch_reference``) are pure fusions: the fill-bit mask, concat and 9-way site mean
are one pass over ``xp`` (dl_fillbit_pool), ``norm(v + mhla(v))`` is one kernel pair, and the
``(B,2304,641)`` concat handed to the SSL head is only materialised when ``lazy_ssl_concat`` is off
(``SSL.forward`` accepts both forms).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functions as Fn
from . import kernels as K
from .config import get_cfg_defaults, get_model_defaults
from .modules import (MLP, CrossModality, FeedForwardLayer, GuidedCrossAttention, MolecularGCN,
                      MultiHeadLinearAttention, PairedMultimodelAttention, ProteinCNN)
from .params import FlatParams
from .ssl import SSL

CONFIGS = {'LAMP': get_model_defaults}


class DrugLAMPBase(nn.Module):
    def __init__(self, n_drug_feature, n_prot_feature, n_hidden=128, **cfg):
        super().__init__()
        if not cfg:
            cfg = get_cfg_defaults()
        drug_padding = cfg["DRUG"]["PADDING"]
        drug_in_feats = cfg["DRUG"]["NODE_IN_FEATS"]
        self.site_len = cfg['PROTEIN']['SITE_LEN']
        self.seq_len_q = cfg['PROTEIN']['SEQ_LEN']
        protein_padding = cfg["PROTEIN"]["PADDING"]
        protein_kernel_size = cfg["PROTEIN"]["KERNEL_SIZE"]
        mlp_in_dim = cfg["DECODER"]["IN_DIM"]
        mlp_binary = cfg["DECODER"]["BINARY"]
        mlp_out_dim = cfg["DECODER"]["OUT_DIM"]
        mlp_hidden_dim = cfg["DECODER"]["HIDDEN_DIM"]
        self.n_drug_feature, self.n_prot_feature = n_drug_feature, n_prot_feature

        self.drug_extractor = MolecularGCN(in_feats=drug_in_feats, dim_embedding=n_hidden,
                                           padding=drug_padding, hidden_feats=[n_hidden] * 3)
        self.protein_extractor = ProteinCNN(n_hidden, [n_hidden] * 3, protein_kernel_size, protein_padding)
        self.ssl_model = SSL(prot_extractor=self.protein_extractor, n_prot_feature=n_prot_feature,
                             drug_ssl_type='simsiam', n_hidden=n_hidden)
        self.cm_model = CrossModality(use_cm=True, hidden_size=n_hidden,
                                      max_margin=cfg["RS"]["MAX_MARGIN"], n_re=cfg["RS"]["RESET_EPOCH"])
        model_cfg = CONFIGS['LAMP'](n_hidden)

        self.lin_d1 = nn.Linear(n_drug_feature + 1, 2 * n_hidden)
        self.act_d = nn.GELU()
        self.d_norm = nn.LayerNorm(2 * n_hidden)
        self.lin_d2 = nn.Linear(2 * n_hidden, n_hidden)

        self.p_adaptor_wo_skip_connect = FeedForwardLayer(n_prot_feature + 1, n_hidden)
        self.lin_p1 = nn.Linear(n_prot_feature + 1, 2 * n_hidden)
        self.act_p = nn.GELU()
        self.p_norm = nn.LayerNorm(2 * n_hidden)
        self.lin_p2 = nn.Linear(2 * n_hidden, n_hidden)

        self.v_gca = GuidedCrossAttention(embed_dim=n_hidden, num_heads=1)
        self.v_mhla = MultiHeadLinearAttention(d_model=n_hidden * 2, d_diff=n_hidden * 8, nhead=8,
                                               dropout=model_cfg.mlha_dropout, activation='gelu')
        self.v_gca_norm = nn.LayerNorm(n_hidden * 2)
        self.x_gca = GuidedCrossAttention(embed_dim=n_hidden, num_heads=1)
        self.x_mhla = MultiHeadLinearAttention(d_model=n_hidden * 2, d_diff=n_hidden * 8, nhead=8,
                                               dropout=model_cfg.mlha_dropout, activation='gelu')
        self.x_gca_norm = nn.LayerNorm(n_hidden * 2)

        self.pmma = PairedMultimodelAttention(config=model_cfg, vis=False)
        self.mlp_classifier = MLP(mlp_in_dim * 2, mlp_hidden_dim * 2, mlp_out_dim * 2, binary=mlp_binary)

        self.lazy_ssl_concat = True
        self.A_v_gca = self.A_x_gca = self.attn = self.guide_attn = None
        self._flat = None

    # ---- reference accessors (basic_model.py:123-132) -----------------------------------------
    def get_cross_attn_mat(self, modality='v'):
        if modality == 'v':
            self.A_v_gca = self.A_v_gca.cpu()
            return self.A_v_gca
        self.A_x_gca = self.A_x_gca.cpu()
        return self.A_x_gca

    def get_inter_attn_mat(self):
        return self.attn, self.guide_attn

    # ---- B200 extras ----------------------------------------------------------------------------
    def flatten_parameters(self) -> FlatParams:
        """Re-home all parameters in one flat buffer (call after ``.cuda()``): one cast kernel
        refreshes every bf16 shadow and one all-reduce covers every gradient."""
        groups = []
        for m in self.modules():
            if hasattr(m, "fused_parameter_groups"):
                groups += m.fused_parameter_groups()
        pmma_ids = {id(p) for p in self.pmma.parameters()}
        groups = [g for g in groups if all(id(p) in pmma_ids for p in g)] + \
                 [g for g in groups if not all(id(p) in pmma_ids for p in g)]
        self._flat = FlatParams(self, groups=groups, first=list(self.pmma.parameters()))
        return self._flat

    def _watch_pmma_inputs(self, *tensors):
        """TrainStep's overlap hook: call `_pmma_grads_ready` once the gradients of PMMA's inputs have
        been computed, i.e. when every PMMA (and decoder-head) parameter gradient is final."""
        cb = getattr(self, "_pmma_grads_ready", None)
        if cb is None or not torch.is_grad_enabled():
            return
        uniq = [t for i, t in enumerate(tensors) if t.requires_grad and all(t is not u for u in tensors[:i])]
        pending = [len(uniq)]

        def hook(g):
            pending[0] -= 1
            if pending[0] == 0:
                cb()
            return g
        for t in uniq:
            t.register_hook(hook)

    # ---- shared pieces of the three forwards ------------------------------------------------------
    def _protein_branch(self, vp, fill_bit_p):
        from .modules import ProteinCNN
        if isinstance(self.protein_extractor, ProteinCNN):              # site mean inside the CNN tail
            return self.protein_extractor(vp, fill_bit_p, site_len=self.site_len)
        v = self.protein_extractor(vp, fill_bit_p)                      # (B, 2304, 128)
        return Fn.SitePoolFn.apply(v, self.site_len)                    # (B, 256, 128)   DrugLAMP.py:35-37

    def _llm_adaptors(self, xp_pool, xd_cat):
        xp = self.p_adaptor_wo_skip_connect(xp_pool, residual=xp_pool)   # DrugLAMP.py:43-45
        xp = Fn.linear(xp, self.lin_p1.weight, self.lin_p1.bias, K.ACT_GELU)
        xp = Fn.layer_norm(xp, self.p_norm.weight, self.p_norm.bias, self.p_norm.eps)
        xp = Fn.linear(xp, self.lin_p2.weight, self.lin_p2.bias)
        xd = Fn.linear(xd_cat, self.lin_d1.weight, self.lin_d1.bias, K.ACT_GELU)     # DrugLAMP.py:50-52
        xd = Fn.layer_norm(xd, self.d_norm.weight, self.d_norm.bias, self.d_norm.eps)
        xd = Fn.linear(xd, self.lin_d2.weight, self.lin_d2.bias)
        return xp, xd

    @staticmethod
    def _guided(gca, mhla, norm, p, d):
        """PGCA -> concat -> MHLA gate + residual + LayerNorm (DrugLAMP.py:55-71)."""
        m, A = gca(p.permute(1, 0, 2), d.permute(1, 0, 2), d.permute(1, 0, 2))
        m = Fn.cat_last(p.to(m.dtype), m.permute(1, 0, 2))
        return mhla.forward_residual_norm(m, norm), A

    def _head(self, f):
        from . import modules as M
        B, Lr, C_ = f.shape
        if Lr % 16 == 0 and Lr > 16:
            f = Fn.SitePoolFn.apply(f, Lr // 16)            # first stage of torch.mean(f, dim=1), compute dtype
        with K.local_compute_dtype(torch.float32 if M.HEAD_PRECISION == "fp32" else None):
            f = Fn.SitePoolFn.apply(f, f.shape[1]).view(B, C_)                      # second stage, head precision
        return self.mlp_classifier(f)

    def _masks(self, xd, xp, need_xd=True):
        # the pooled / concatenated LLM features feed 641- and 385-input linears: they are written
        # with zero columns up to the TMA alignment so no padding copy is needed downstream
        al = Fn._align()
        bit_p, xp_cat, xp_pool = K.fillbit_pool(xp, self.site_len, want_cat=not self.lazy_ssl_concat, pad_to=al)
        bit_d = xd_cat = xd_lin = None
        if need_xd:
            bit_d, xd_cat, xd_lin = K.fillbit_pool(xd, 1, want_cat=True, want_pooled=True, pad_to=al)
        return bit_p, xp_cat, xp_pool, bit_d, xd_cat, xd_lin

    def forward(self, vd, vp, xd, xp, mode="train"):
        raise NotImplementedError


class DrugLAMP(DrugLAMPBase):
    """reference ``model/DrugLAMP.py:8-79``."""

    def _forward(self, vd, vp, xd, xp):
        if self._flat is not None:
            self._flat.sync()
        side = Fn.branch_stream(xp)
        self._branch_streams = [] if side is None else [side]      # what the last forward used besides the caller's
        if side is None:
            vd = self.drug_extractor(vd)                                    # (B, 512, 128)
            bit_p, xp_cat, xp_pool, _, xd_cat, xd_lin = self._masks(xd, xp)
            vpf = self._protein_branch(vp, bit_p)
            xpa, xda = self._llm_adaptors(xp_pool, xd_lin)
            mv, self.A_v_gca = self._guided(self.v_gca, self.v_mhla, self.v_gca_norm, vpf, vd)
            mx, self.A_x_gca = self._guided(self.x_gca, self.x_mhla, self.x_gca_norm, xpa, xda)
        else:
            # Two branches that only meet at PMMA (DrugLAMP.py:55-72) run on two streams:
            #   main  fill bit + site means of xp (one pass over 5.9 MB per pair, HBM bound) -> ProteinCNN
            #         (tensor bound) -> v_gca / v_mhla
            #   side  xd fill bit -> LLM adaptors -> x_gca / x_mhla
            #   gcn   MolecularGCN (a few thousand rows: launch-latency bound, leaves the SMs idle)
            # autograd runs every backward node on its forward stream, so the backward forks the same way.
            main = torch.cuda.current_stream()
            gcn = Fn.branch_stream(xp, 1)          # its own stream: its backward (a serial chain of tiny
            self._branch_streams.append(gcn)       # kernels) must not queue up behind the x branch's
            gcn.wait_stream(main)
            with torch.cuda.stream(gcn):
                vd = self.drug_extractor(vd)
                ev_vd = gcn.record_event()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                al = Fn._align()
                _, xd_cat, xd_lin = K.fillbit_pool(xd, 1, want_cat=True, want_pooled=True, pad_to=al)
            bit_p, xp_cat, xp_pool = K.fillbit_pool(xp, self.site_len, want_cat=not self.lazy_ssl_concat, pad_to=al)
            ev_fill = main.record_event()
            with torch.cuda.stream(side):
                side.wait_event(ev_fill)
                Fn.crosses(side, xp_pool)
                xpa, xda = self._llm_adaptors(xp_pool, xd_lin)
                mx, self.A_x_gca = self._guided(self.x_gca, self.x_mhla, self.x_gca_norm, xpa, xda)
            vpf = self._protein_branch(vp, bit_p)
            main.wait_event(ev_vd)
            Fn.crosses(main, vd)
            mv, self.A_v_gca = self._guided(self.v_gca, self.v_mhla, self.v_gca_norm, vpf, vd)
            main.wait_stream(side)
            Fn.crosses(main, mx, self.A_x_gca, xpa, xda, xd_cat)
        ssl = {'vp': vp, 'xp': xp if xp_cat is None else xp_cat, 'fill_bit_p': bit_p, 'vd': vd, 'xd': xd_cat}
        self._watch_pmma_inputs(mx, mv)
        f, self.attn, self.guide_attn = self.pmma(mx, mv)
        score = self._head(f)
        return vd, vpf, xda, xpa, ssl, score

    def forward(self, vd, vp, xd, xp, mode="train"):
        vd, vp, _, _, ssl, score = self._forward(vd, vp, xd, xp)
        if mode == "train":
            return vd, vp, ssl, None, score
        elif mode == "eval":
            return vd, vp, score, self.attn


class DrugLAMP2C2P(DrugLAMP):
    """reference ``model/DrugLAMP2C2P.py:8-90``: DrugLAMP plus the 2C2P inputs in slot 4."""

    def forward(self, vd, vp, xd, xp, mode="train"):
        vd, vp, xda, xpa, ssl, score = self._forward(vd, vp, xd, xp)
        cp = {'prot': vp, 'aug_prot': xpa, 'drug': vd, 'aug_drug': xda}
        if mode == "train":
            return vd, vp, ssl, cp, score
        elif mode == "eval":
            return vd, vp, score, self.attn


class DrugLAMPwoLLM(DrugLAMPBase):
    """reference ``model/DrugLAMPwoLLM.py:8-52``: no LLM branch, ``pmma(mv, mv)``; the fill bit is
    still derived from ``xp`` (``:11-13``)."""

    def forward(self, vd, vp, xd, xp, mode="train"):
        if self._flat is not None:
            self._flat.sync()
        vd = self.drug_extractor(vd)
        bit_p, _, _ = K.fillbit_pool(xp, self.site_len, want_cat=False, want_pooled=False)
        ssl = {'vp': vp, 'xp': None, 'fill_bit_p': bit_p, 'vd': vd, 'xd': None, 'p_mode': 'vp'}
        vpf = self._protein_branch(vp, bit_p)
        mv, self.A_v_gca = self._guided(self.v_gca, self.v_mhla, self.v_gca_norm, vpf, vd)
        self._watch_pmma_inputs(mv)
        f, self.attn, self.guide_attn = self.pmma(mv, mv)
        score = self._head(f)
        if mode == "train":
            return vd, vpf, ssl, None, score
        elif mode == "eval":
            return vd, vpf, score, self.attn
