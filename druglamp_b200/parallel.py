"""Data-parallel pieces of the hot path (one process per GPU, ``torch.distributed`` over NCCL).

The reference scales only through Lightning DDP (``trainer.py:147``) and never exchanges embeddings
(SURVEY 2.4, 8e).  Two exchange steps exist here:

* gradient all-reduce -- ONE sum all-reduce of the flat gradient buffer per step
  (``train.TrainStep._reduce``; the reference issues up to three bucketed all-reduces per step);
* the CrossModality all-gather that gives the 2C2P triplet loss its *global* negatives
  (BASELINE.json configs[2]): every rank contributes the pooled features of its local pairs, ids and
  labels; every rank then rebuilds the global label matrix exactly as a single process would on the
  concatenated batch (dedup "last index wins" over the rank-ordered list), runs Mean2Embed /
  latents / triplet loss on the global set, and back-propagates into its own slice.

Equivalence pinned by tests/test_parallel_gloo_cpu.py: N ranks == one process on the concatenated
batch, for the loss and for the gradients reaching each rank's local features.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch.autograd import Function


class AllGatherRows(Function):
    """Concatenate `x` (rows differ per rank allowed) from all ranks along dim 0, in rank order.

    Backward assumes the downstream computation is replicated on every rank and that parameter
    gradients are later *averaged* over ranks (DDP convention): each rank keeps the slice of the
    incoming gradient that belongs to its rows, multiplied by world_size, so that after averaging
    the backbone sees the same gradient a single process would."""

    @staticmethod
    def forward(ctx, x, group, average_downstream):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        n_local = torch.tensor([x.shape[0]], device=x.device, dtype=torch.int64)
        sizes = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(sizes, n_local, group=group)
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes)
        pad = x if x.shape[0] == mx else torch.cat([x, x.new_zeros((mx - x.shape[0],) + x.shape[1:])])
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad.contiguous(), group=group)
        ctx.meta = (sizes, rank, world, average_downstream)
        return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)

    @staticmethod
    def backward(ctx, g):
        sizes, rank, world, average_downstream = ctx.meta
        start = sum(sizes[:rank])
        gs = g[start:start + sizes[rank]]
        return (gs * world if average_downstream else gs), None, None


def all_gather_rows(x: torch.Tensor, group=None, average_downstream: bool = True) -> torch.Tensor:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return AllGatherRows.apply(x, group, average_downstream)


def all_gather_meta(meta: List[dict], group=None) -> List[dict]:
    """Rank-ordered concatenation of the per-pair metadata (ids and labels are tiny: host objects)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(meta)
    out: List[Optional[List[dict]]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, [{"Prot_ID": m["Prot_ID"], "Drug_ID": m["Drug_ID"], "Y": int(m["Y"])} for m in meta],
                           group=group)
    return [m for part in out for m in part]


def global_cross_modality_loss(cm, prot, aug_prot, drug, aug_drug, meta, group=None, pool_fn=None,
                               average_downstream: bool = True):
    """CrossModality loss with all-gathered (global) negatives.

    `cm` is a ``modules.CrossModality``.  `pool_fn(seq) -> (B_local, hidden)` is the mean over the
    sequence axis (defaults to the dl_site_pool kernel); the pooled features -- 4 x (B_local, 128)
    values per rank instead of the (B, L, 128) sequences -- are what crosses NVLink.
    Single-process equivalence: identical to ``cm(prot, aug_prot, drug, aug_drug, meta)`` evaluated
    on the rank-ordered concatenation of all ranks' batches."""
    from . import functions as Fn
    if pool_fn is None:
        def pool_fn(seq):
            return Fn.SitePoolFn.apply(seq, seq.shape[1]).view(seq.shape[0], seq.shape[2])
    pooled = [all_gather_rows(pool_fn(t), group, average_downstream) for t in (prot, aug_prot, drug, aug_drug)]
    targets = cm.prepare(all_gather_meta(meta, group))
    if targets.G.device != pooled[0].device:
        targets = targets.to(pooled[0].device)
    pl, dl = cm.latents_from_pooled(*pooled, targets)
    return cm.loss_from_latents(pl, dl, targets.G)
