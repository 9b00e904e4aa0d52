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

from typing import List

import torch
import torch.distributed as dist
from torch.autograd import Function


class AllGatherRows(Function):
    """Concatenate `x` from all ranks along dim 0, in rank order.  Every rank contributes the SAME
    number of rows (data-parallel shards of one global batch), so this is ONE fixed-size
    ``all_gather_into_tensor`` (ncclAllGather into a preallocated buffer): no size exchange, no host
    synchronisation, capturable in a CUDA graph.

    Backward assumes the downstream computation is replicated on every rank and that parameter
    gradients are later *averaged* over ranks (DDP convention): each rank keeps the slice of the
    incoming gradient that belongs to its rows, multiplied by world_size, so that after averaging
    the backbone sees the same gradient a single process would."""

    @staticmethod
    def forward(ctx, x, group, average_downstream):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
        ctx.meta = (x.shape[0], rank, world, average_downstream)
        return out

    @staticmethod
    def backward(ctx, g):
        n, rank, world, average_downstream = ctx.meta
        gs = g[rank * n:(rank + 1) * n]
        return (gs * world if average_downstream else gs), None, None


def all_gather_rows(x: torch.Tensor, group=None, average_downstream: bool = True) -> torch.Tensor:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return AllGatherRows.apply(x, group, average_downstream)


def _id64(s) -> int:
    """Stable 63-bit hash of an entity id (str or int): ids travel as int64, not as pickled objects."""
    import hashlib
    return int.from_bytes(hashlib.blake2b(str(s).encode(), digest_size=8).digest(), "little") >> 1


def all_gather_meta(meta: List[dict], group=None, device=None) -> List[dict]:
    """Rank-ordered concatenation of the per-pair metadata as ONE fixed-size tensor collective: each
    pair contributes (hash64(Prot_ID), hash64(Drug_ID), Y) as an int64 row.  The label matrix only
    needs id EQUALITY, which the hashes preserve.  (When every rank can enumerate the whole global
    batch itself -- a DistributedSampler with a shared seed does -- no exchange is needed at all:
    pass the global meta list to ``CrossModality.prepare`` directly; bench.py --config 2c2p does.)"""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(meta)
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    loc = torch.tensor([[_id64(m["Prot_ID"]), _id64(m["Drug_ID"]), int(m["Y"])] for m in meta],
                       dtype=torch.int64, device=device)
    out = torch.empty((world * loc.shape[0], 3), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, loc, group=group)
    return [{"Prot_ID": p, "Drug_ID": d, "Y": y} for p, d, y in out.cpu().tolist()]


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
            return Fn.seq_mean(seq)
    pooled = [all_gather_rows(pool_fn(t), group, average_downstream) for t in (prot, aug_prot, drug, aug_drug)]
    targets = cm.prepare(all_gather_meta(meta, group))
    if targets.G.device != pooled[0].device:
        targets = targets.to(pooled[0].device)
    pl, dl = cm.latents_from_pooled(*pooled, targets)
    return cm.loss_from_latents(pl, dl, targets.G)
