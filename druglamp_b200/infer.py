"""A CUDA-graph-captured eval-mode forward over the sm_100a hot path (forward-only scoring).

The reference scores pairs in ``ExpModule.validation_step`` / ``test_step`` (``trainer.py:256-292``):
``model(vd, vp, xd, xp, mode='eval')`` with BatchNorm on its running statistics and dropout off, then
``binary_cross_entropy(score, labels)`` for the probabilities ``n`` and the loss.  This class is the
B200-side equivalent for BASELINE.json configs[4] (batch-1024 inference sweep, 128 pairs per GPU,
data-parallel replicas with no exchange step): one graph replay per batch, no autograd bookkeeping,
no Python between the ~120 launches.
"""
from __future__ import annotations

import weakref

import torch

from . import _lib as L
from .functions import forward_only
from .modules import binary_cross_entropy
from .train import StaticBatch


class InferStep:
    def __init__(self, model):
        self.model = model.eval()
        self.flat = model._flat or model.flatten_parameters()
        self._graphs = weakref.WeakKeyDictionary()      # StaticBatch -> graph: dies with the batch
        self._pool = None
        self.launches_per_step = 0

    def eager(self, sb: StaticBatch):
        """(probabilities (B,), loss) of one batch; tensors are overwritten by the next call on the
        same batch when it is replayed from a graph."""
        with torch.no_grad(), forward_only():
            out = self.model(*sb.model_inputs(), mode="eval")
            n, loss = binary_cross_entropy(out[2], sb.y)
        return n, loss

    def capture(self, sb: StaticBatch, warmup: int = 2) -> None:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.eager(sb)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(g, pool=self._pool):
            n, loss = self.eager(sb)
        if self._pool is None:
            self._pool = g.pool()
        self.launches_per_step = L.launch_count() - n0
        self._graphs[sb] = (g, n, loss)

    def replay(self, sb: StaticBatch):
        g, n, loss = self._graphs[sb]
        g.replay()
        return n, loss
