"""A CUDA-graph-captured data-parallel training step over the sm_100a hot path.

One step = zero the flat gradient buffer -> model forward -> ``binary_cross_entropy`` -> backward
-> (data-parallel: one NCCL sum all-reduce of the flat gradient buffer) -> fused AdamW over the
flat parameter buffer (which also refreshes the bf16 shadows).  Everything on the device side is a
hand-written kernel launch (a few PyTorch tensor adds / copies remain as autograd glue); capturing it
in a CUDA graph removes the Python/launch overhead that otherwise dominates at batch 64.

The reference does this through Lightning's manual optimisation with up to three backward passes
and three all-reduces per step (``trainer.py:179-231``); this class is the B200-side
equivalent of the classification step only and is what ``bench.py`` times.
"""
from __future__ import annotations

import weakref
from typing import List

import torch

from . import _lib as L
from . import functions as Fn
from .graph import BatchedMolGraph
from .modules import binary_cross_entropy
from .params import FlatAdamW


class StaticBatch:
    """Device-resident input buffers with fixed addresses (what a captured graph reads)."""

    def __init__(self, batch, device):
        self.graph: BatchedMolGraph = batch.graph.to(device)
        self.h = self.graph.ndata["h"]
        self.vp = batch.vp.to(device)
        self.xd = batch.xd.to(device)
        self.xp = batch.xp.to(device)
        self.y = batch.y.to(device).float()
        self.n_pairs = int(batch.y.shape[0])
        self._stage = {}                          # device staging of packed rows (load_from_packed)

    def tensors(self) -> List[torch.Tensor]:
        """What crosses PCIe for one batch in the reference collate's layout: node features, tokens,
        the dense LLM embeddings, labels and the batched graph's raw edge list (the CSR carrier is
        rebuilt from it on the device, ``BatchedMolGraph.rebuild_``)."""
        g = self.graph
        c = g.compact()
        return [self.h, self.vp, self.xd, self.xp, self.y, g.src, g.dst] + ([] if c is None else c.tensors())

    def host_copy(self, pin=True) -> List[torch.Tensor]:
        out = []
        for t in self.tensors():
            c = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=pin)
            c.copy_(t)
            out.append(c)
        return out

    def load_from(self, host: List[torch.Tensor]) -> int:
        """Asynchronous H2D refresh of every input buffer, then the device-side CSR construction
        from the edge list (dl_csr_build: no host sync); returns the bytes copied."""
        n = 0
        for dst, src in zip(self.tensors(), host):
            dst.copy_(src, non_blocking=True)
            n += src.numel() * src.element_size()
        self.graph.rebuild_()
        return n

    # ---- packed wire format (druglamp_b200.collate): untiled embedding rows over PCIe ---------
    def _small(self) -> List[torch.Tensor]:
        g = self.graph
        c = g.compact()
        return [self.h, self.vp, self.y, g.src, g.dst] + ([] if c is None else c.tensors())

    def host_copy_packed(self, batch, pin=True) -> dict:
        """Host image of this batch with the LLM embeddings as packed rows (what the dataset yields
        before ``utils.tail_pad`` / ``repeat_pad``) instead of the dense padded tensors."""
        from .collate import pack_rows
        small = []
        for t in self._small():
            c = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=pin)
            c.copy_(t)
            small.append(c)
        d_blocks, p_blocks = batch.llm_blocks()
        return {"small": small,
                "xd": pack_rows(d_blocks, self.xd.shape[1], repeat=False, pin=pin),
                "xp": pack_rows(p_blocks, self.xp.shape[1], repeat=True, pin=pin)}

    def load_from_packed(self, host: dict) -> int:
        """Asynchronous H2D of the packed image, then the device-side padding / tiling into the
        dense input buffers (bit-identical to the reference collate's tensors); returns bytes copied."""
        from . import kernels as K
        n = 0
        for dst, src in zip(self._small(), host["small"]):
            dst.copy_(src, non_blocking=True)
            n += src.numel() * src.element_size()
        self.graph.rebuild_()
        for key, dense in (("xd", self.xd), ("xp", self.xp)):
            pk = host[key]
            stage = self._stage.get(key)
            if stage is None or stage[0].shape[0] < pk.rows.shape[0]:
                cap = max(pk.rows.shape[0], 0 if stage is None else stage[0].shape[0])
                stage = (torch.empty((cap, pk.rows.shape[1]), dtype=torch.float32, device=dense.device),
                         torch.empty(pk.offsets.shape, dtype=torch.int32, device=dense.device))
                self._stage[key] = stage
            rows = stage[0][:pk.rows.shape[0]]
            rows.copy_(pk.rows, non_blocking=True)
            stage[1].copy_(pk.offsets, non_blocking=True)
            K.expand_rows(rows, stage[1], dense, pk.repeat)
            n += pk.nbytes()
        return n

    def model_inputs(self):
        self.graph.ndata["h"] = self.h            # MolecularGCN pops it every forward
        return self.graph, self.vp, self.xd, self.xp


class TrainStep:
    def __init__(self, model, lr=1e-4, weight_decay=1e-2, world_size=1, process_group=None):
        self.model = model
        self.flat = model._flat or model.flatten_parameters()
        self.opt = FlatAdamW(self.flat, lr=lr, weight_decay=weight_decay)
        self.world_size = world_size
        self.pg = process_group
        if world_size > 1 and torch.distributed.is_initialized():
            from .functions import set_seed_stream
            set_seed_stream(torch.distributed.get_rank(process_group))
        # Overlapped with the backward: the PMMA parameters (73 % of the model, laid out first in the flat
        # buffer) have final gradients as soon as the backward reaches PMMA's inputs.  From that point an
        # auxiliary stream (a) all-reduces their range (N > 1) and (b) runs their AdamW update, while MHLA /
        # PGCA / CNN / GCN still run; the rest of the buffer follows the backward.  The collectives and the
        # early update are part of the captured step graph.
        import os
        split = self.flat.head_numel < self.flat.numel
        self.overlap = world_size > 1 and split and os.environ.get("DL_NO_OVERLAP", "0") == "0"
        # (b) is off by default: measured on one B200 it is neutral (4.633 vs 4.622 ms) -- the update is
        # HBM bound and only competes with the backward kernels it overlaps; DL_EARLY_UPDATE=1 turns it on
        self.early_update = (split and (world_size == 1 or self.overlap)
                             and os.environ.get("DL_EARLY_UPDATE", "0") == "1")
        self._aux = torch.cuda.Stream()
        self._comm = self._aux
        self._joined = True
        self._head_updated = False
        if self.overlap or self.early_update:
            model._pmma_grads_ready = self._pmma_ready
        self.loss = torch.zeros((), dtype=torch.float32, device=self.flat.flat.device)
        self._graphs = weakref.WeakKeyDictionary()      # StaticBatch -> (graph, graph): dies with the batch
        self._pool = None
        self.launches_per_step = 0

    # ---- pieces ---------------------------------------------------------------------------------
    def _fwd_bwd(self, sb: StaticBatch) -> None:
        from . import kernels as K
        main = torch.cuda.current_stream()
        # the gradient buffer is cleared beside the forward (nothing reads it before the backward)
        self._aux.wait_stream(main)
        with torch.cuda.stream(self._aux):
            self.flat.zero_grad()
        # dropout seeds advance with the optimiser's device-side step counter, so a replayed graph
        # draws fresh masks every step (forward and backward of one step read the same value: the
        # counter moves in _update)
        K.set_dropout_step(self.opt.step_count)
        try:
            out = self.model(*sb.model_inputs())
            _, loss = binary_cross_entropy(out[4], sb.y)
            main.wait_stream(self._aux)
            # weight-gradient GEMMs leave the dX chain for side streams; joined when the context closes
            with Fn.deferred_weight_grads():
                loss.backward()
        finally:
            K.set_dropout_step(None)
        self.loss.copy_(loss.detach())
        if self.overlap:
            if self._joined:            # the hook never fired (no PMMA gradient): reduce everything here
                torch.distributed.all_reduce(self.flat.grad, group=self.pg)
            else:
                torch.distributed.all_reduce(self.flat.grad[self.flat.head_numel:], group=self.pg)
        if not self._joined:
            main.wait_stream(self._aux)
            self._joined = True

    def _pmma_ready(self) -> None:
        """Called from inside the backward (models._watch_pmma_inputs): PMMA's gradients are final."""
        aux = self._aux
        aux.wait_stream(torch.cuda.current_stream())
        for s in getattr(self.model, "_branch_streams", []):    # the hook may fire on any branch stream
            aux.wait_stream(s)
        if Fn._wgrad_defer is not None:                          # PMMA's deferred weight gradients
            Fn._wgrad_defer.join(aux)
        with torch.cuda.stream(aux):
            if self.overlap:
                torch.distributed.all_reduce(self.flat.grad[:self.flat.head_numel], group=self.pg)
            if self.early_update:
                # the counter is ticked by the final range in _update: dropout kernels of the remaining
                # backward still read this step's value
                self.opt.step(grad_scale=1.0 / self.world_size, lo=0, hi=self.flat.head_numel, tick=False)
                self._head_updated = True
        self._joined = False

    def _reduce(self) -> None:
        if self.world_size > 1 and not self.overlap:
            torch.distributed.all_reduce(self.flat.grad, group=self.pg)

    def _update(self) -> None:
        lo = self.flat.head_numel if self._head_updated else 0
        self._head_updated = False
        self.opt.step(grad_scale=1.0 / self.world_size, lo=lo)

    def eager(self, sb: StaticBatch) -> torch.Tensor:
        self._fwd_bwd(sb)
        self._reduce()
        self._update()
        return self.loss

    # ---- CUDA graphs ------------------------------------------------------------------------------
    def capture(self, sb: StaticBatch, warmup: int = 2) -> None:
        """Capture fwd+bwd and the optimizer update for this batch's buffers (the NCCL all-reduce
        between them stays a stream-ordered eager call)."""
        # The warm-up steps (lazy initialisation, allocator pool) and the capture itself must not
        # train on the capture batch: parameters, Adam moments, the step counter and every module
        # buffer (BatchNorm running statistics) are restored afterwards.
        opt = self.opt
        keep = [t.clone() for t in (self.flat.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_count)]
        bufs = [(b, b.clone()) for b in self.model.buffers()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.eager(sb)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(g1, pool=self._pool):
            self._fwd_bwd(sb)
        if self._pool is None:
            self._pool = g1.pool()
        opt.refresh_active()          # which parameters this step's backward reaches (host -> device, not captured)
        with torch.cuda.graph(g2, pool=self._pool):
            self._update()
        self.launches_per_step = L.launch_count() - n0
        self._graphs[sb] = (g1, g2)
        with torch.no_grad():
            for dst, src in zip((self.flat.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_count), keep):
                dst.copy_(src)
            for b, c in bufs:
                b.copy_(c)
        self.flat.sync(force=True)

    def replay(self, sb: StaticBatch) -> torch.Tensor:
        g1, g2 = self._graphs[sb]
        g1.replay()
        self._reduce()
        g2.replay()
        return self.loss
