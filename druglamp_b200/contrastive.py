"""DrugLAMP2C2P training step with GLOBAL in-batch negatives (BASELINE.json configs[2]: contrastive
pretraining, global batch 4096 = 4096 / G pairs per GPU, NCCL all-gathered negatives).

The reference computes the 2C2P loss on the local batch only (``trainer.py:210-223`` calls
``cm_model(**cm_input, meta)`` with what ``model/DrugLAMP2C2P.py:54-63`` hands back, and Lightning DDP
never exchanges embeddings).  Here every pair of the global batch is a negative for every
other: what crosses NVLink is the four pooled feature rows per pair (4 x 128 values), never a sequence.

A rank's share of the global batch does not fit one forward (4096 pairs of inputs are 28 GB), so it is
stepped as micro-batches with gradient accumulation; the contrastive loss couples ALL pairs, hence two
passes (the "gradient cache" scheme):

  1. features   every local micro-batch through the extractors and adaptors only (MolecularGCN,
                ProteinCNN + site pooling, LLM adaptors -- the only layers the 2C2P inputs depend on;
                no autograd, BatchNorm buffers untouched) -> pooled (4, B_local, 128)
  2. gather     ONE fixed-size all_gather_into_tensor of the pooled rows -> (4, N, 128)
  3. contrast   CrossModality on the global set (replicated on every rank: 8.6 GFLOP at N = 4096):
                loss, its parameter gradients, and d loss / d pooled for the local rows
  4. backbone   every local micro-batch again, full forward: classification loss / n_micro plus the
                inner product <pooled, d pooled> -- whose gradient is exactly the contrastive loss's
                -- backward, accumulating into the flat gradient buffer
  5. one gradient all-reduce, AdamW.

The label matrix needs the ids of the whole global batch on every rank.  A distributed sampler with a
shared seed lets every rank enumerate them without communication (what bench.py does); otherwise
``parallel.all_gather_meta`` moves them as hashed int64 rows in one collective.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import _lib as L
from . import functions as Fn
from . import kernels as K
from .modules import CMTargets, binary_cross_entropy
from .params import FlatAdamW
from .train import StaticBatch


def _pool(seq: torch.Tensor) -> torch.Tensor:
    """mean over the sequence axis (``cross_modality.py:151-154``) -> (B, hidden), fp32."""
    return Fn.seq_mean(seq).float()


class ContrastiveStep:
    def __init__(self, model, targets: CMTargets, n_local: int, lr=1e-4, weight_decay=1e-2, cm_weight=1.0,
                 world_size=1, rank=0, process_group=None):
        """targets: ``CrossModality.prepare`` of the GLOBAL meta list (rank-ordered);
        n_local: pairs of the global batch owned by this rank (its rows are [rank*n_local, +n_local))."""
        self.model = model
        self.flat = model._flat or model.flatten_parameters()
        self.opt = FlatAdamW(self.flat, lr=lr, weight_decay=weight_decay)
        self.world_size, self.rank, self.pg = world_size, rank, process_group
        dev = self.flat.flat.device
        self.targets = targets.to(dev)
        self.n_local = n_local
        self.cm_weight = cm_weight
        self.pooled = torch.zeros((4, n_local, model.cm_model.to_prot_latent.in_features // 2),
                                  dtype=torch.float32, device=dev)
        self.pooled_all = torch.zeros((4, n_local * world_size, self.pooled.shape[2]), dtype=torch.float32, device=dev)
        self.dpooled = torch.zeros_like(self.pooled)
        self.cls_loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.cm_loss = torch.zeros((), dtype=torch.float32, device=dev)
        self._g_feat, self._g_back, self._g_cm, self._g_opt = {}, {}, None, None
        self._pool = None
        self.launches = 0

    # ---- pass 1: features ---------------------------------------------------------------------------
    def _features(self, sb: StaticBatch, out: torch.Tensor) -> None:
        """out (4, B, 128) <- pooled (prot, aug_prot, drug, aug_drug) of one micro-batch."""
        m = self.model
        with torch.no_grad(), Fn.frozen_bn_buffers(), Fn.forward_only():
            g, vp, xd, xp = sb.model_inputs()
            vd = m.drug_extractor(g)
            bit_p, _, xp_pool, _, _, xd_lin = m._masks(xd, xp)
            vpf = m._protein_branch(vp, bit_p)
            xpa, xda = m._llm_adaptors(xp_pool, xd_lin)
            for i, t in enumerate((vpf, xpa, vd, xda)):
                out[i].copy_(_pool(t))

    # ---- pass 3: the contrastive loss on the gathered rows --------------------------------------------
    def _contrast(self) -> None:
        cm = self.model.cm_model
        leaves = self.pooled_all.detach().clone().requires_grad_(True)
        pl, dl = cm.latents_from_pooled(leaves[0], leaves[1], leaves[2], leaves[3], self.targets)
        loss = cm.loss_from_latents(pl, dl, self.targets.G) * self.cm_weight
        loss.backward()
        self.cm_loss.copy_(loss.detach())
        lo = self.rank * self.n_local
        # parameter gradients are averaged over ranks after the all-reduce: the backbone's share of
        # this (replicated, un-averaged) loss is pre-multiplied by the world size
        self.dpooled.copy_(leaves.grad[:, lo:lo + self.n_local] * float(self.world_size))

    # ---- pass 4: backbone ---------------------------------------------------------------------------------
    def _backbone(self, sb: StaticBatch, dp: torch.Tensor, cls_scale: float) -> None:
        K.set_dropout_step(self.opt.step_count)
        try:
            out = self.model(*sb.model_inputs())
            cp = out[3]
            _, cls = binary_cross_entropy(out[4], sb.y)
            total = cls * cls_scale
            for i, k in enumerate(("prot", "aug_prot", "drug", "aug_drug")):
                total = total + (_pool(cp[k]) * dp[i]).sum()
            with Fn.deferred_weight_grads():
                total.backward()
        finally:
            K.set_dropout_step(None)
        self.cls_loss.add_(cls.detach() * cls_scale)

    # ---- the step ---------------------------------------------------------------------------------------------
    def step(self, micro: Sequence[StaticBatch], graphs: bool = False, update: bool = True) -> torch.Tensor:
        """One optimiser step over this rank's micro-batches (len(micro) * pairs each == n_local).
        update=False stops after the gradient all-reduce (tests read ``flat.grad``)."""
        nm = len(micro)
        B = micro[0].n_pairs
        assert nm * B == self.n_local
        self.flat.zero_grad()
        self.cls_loss.zero_()
        for i, sb in enumerate(micro):
            if graphs:
                self._g_feat[sb].replay()
                self.pooled[:, i * B:(i + 1) * B].copy_(self._feat_buf)
            else:
                self._features(sb, self.pooled[:, i * B:(i + 1) * B])
        if self.world_size > 1:
            # (4, N, 128) with N rank-major: gather per feature kind
            for k in range(4):
                torch.distributed.all_gather_into_tensor(self.pooled_all[k], self.pooled[k].contiguous(), group=self.pg)
        else:
            self.pooled_all.copy_(self.pooled)
        if graphs:
            self._g_cm.replay()
        else:
            self._contrast()
        for i, sb in enumerate(micro):
            if graphs:
                self._dp_buf.copy_(self.dpooled[:, i * B:(i + 1) * B])
                self._g_back[sb].replay()
            else:
                self._backbone(sb, self.dpooled[:, i * B:(i + 1) * B], 1.0 / nm)
        if self.world_size > 1:
            torch.distributed.all_reduce(self.flat.grad, group=self.pg)
        if not update:
            return self.cls_loss + self.cm_loss
        if graphs:
            self._g_opt.replay()
        else:
            self.opt.step(grad_scale=1.0 / self.world_size)
        return self.cls_loss + self.cm_loss

    # ---- CUDA graphs ---------------------------------------------------------------------------------------------
    def capture(self, distinct: List[StaticBatch], n_micro: int) -> None:
        """Capture the per-micro-batch graphs (features, backbone) for each DISTINCT resident batch, the
        contrastive graph and the optimiser graph.  State touched while capturing is restored."""
        opt = self.opt
        keep = [t.clone() for t in (self.flat.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_count)]
        bufs = [(b, b.clone()) for b in self.model.buffers()]
        B = distinct[0].n_pairs
        dev = self.pooled.device
        self._feat_buf = torch.zeros((4, B, self.pooled.shape[2]), dtype=torch.float32, device=dev)
        self._dp_buf = torch.zeros_like(self._feat_buf)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                       # warm-up: lazy initialisation, allocator pool
            self.flat.zero_grad()
            for sb in distinct:
                self._features(sb, self._feat_buf)
                self._backbone(sb, self._dp_buf, 1.0 / n_micro)
            self._contrast()
            self.opt.step(grad_scale=1.0 / self.world_size)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = L.launch_count()
        per_micro = 0
        for sb in distinct:
            gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            a0 = L.launch_count()
            with torch.cuda.graph(gf, pool=self._pool):
                self._features(sb, self._feat_buf)
            if self._pool is None:
                self._pool = gf.pool()
            with torch.cuda.graph(gb, pool=self._pool):
                self._backbone(sb, self._dp_buf, 1.0 / n_micro)
            per_micro = L.launch_count() - a0
            self._g_feat[sb], self._g_back[sb] = gf, gb
        a0 = L.launch_count()
        self._g_cm = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_cm, pool=self._pool):
            self._contrast()
        opt.refresh_active()
        self._g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_opt, pool=self._pool):
            self.opt.step(grad_scale=1.0 / self.world_size)
        self.launches = per_micro * n_micro + (L.launch_count() - a0)
        del n0
        with torch.no_grad():
            for dst, src in zip((self.flat.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_count), keep):
                dst.copy_(src)
            for b, c in bufs:
                b.copy_(c)
        self.flat.sync(force=True)
