"""Graph carrier for the molecular GCN (SURVEY.md section 8b "Graph carrier").

The reference hands ``MolecularGCN.forward`` a batched ``DGLGraph``
(``model/basic_model.py:147-153``; built by ``handler/dataset.py:212-222`` and
``utils.py:326-334``).  The B200 path wants the adjacency as CSR *by destination*
so that one warp owns one destination row of the segment-sum (forward), and CSR
*by source* for the transposed aggregation in backward.  ``BatchedMolGraph``
holds both, plus the symmetric-normalisation vectors
``deg.clamp(min=1) ** -0.5`` of ``GraphConv.forward`` (``basic_model.py:596-603``,
``:623-630``).  Duplicate edges are kept and counted (App. A5: real atoms carry two
self loops).

It duck-types the few DGLGraph members the reference touches (``ndata`` with
``pop``, ``batch_size``, ``num_nodes()``, ``edges()``, ``in_degrees()``,
``out_degrees()``, ``local_scope()``, ``to()``), so the *reference's own*
``MolecularGCN`` never sees it but ``model/DrugLAMP*.py`` can pass it through.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch


def _csr(sort_key: torch.Tensor, other: torch.Tensor, n: int):
    """CSR rows = sort_key; column ids = ``other`` in stable edge order."""
    order = torch.argsort(sort_key, stable=True)
    counts = torch.bincount(sort_key, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int32, device=sort_key.device)
    indptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return indptr, other[order].to(torch.int32).contiguous(), counts


class BatchedMolGraph:
    is_block = False

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int,
                 batch_size: int, h: Optional[torch.Tensor] = None):
        if src.shape != dst.shape or src.dim() != 1:
            raise ValueError("src/dst must be 1-D tensors of equal length")
        self._n = int(num_nodes)
        self.batch_size = int(batch_size)
        self.src = src.long()
        self.dst = dst.long()
        if self.src.numel() and (int(self.src.max()) >= self._n or int(self.dst.max()) >= self._n
                                 or int(self.src.min()) < 0 or int(self.dst.min()) < 0):
            raise ValueError("edge endpoint out of range")
        self.ndata = {} if h is None else {"h": h}
        # CSR by destination (forward aggregation) and by source (backward)
        self.indptr, self.indices, in_counts = _csr(self.dst, self.src, self._n)
        self.indptr_t, self.indices_t, out_counts = _csr(self.src, self.dst, self._n)
        self.in_deg = in_counts
        self.out_deg = out_counts
        self.norm_dst = in_counts.clamp(min=1).to(torch.float32).pow(-0.5)
        self.norm_src = out_counts.clamp(min=1).to(torch.float32).pow(-0.5)
        # evaluated once here so that GraphConv's guard costs no host sync per layer
        self._zero_in = bool((in_counts == 0).any()) if self._n else False

    # ---- DGLGraph duck-typing -------------------------------------------------
    def num_nodes(self) -> int:
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self) -> int:
        return int(self.src.numel())

    def edges(self):
        return self.src, self.dst

    def in_degrees(self):
        return self.in_deg

    def out_degrees(self):
        return self.out_deg

    @contextlib.contextmanager
    def local_scope(self):
        yield

    @property
    def device(self):
        return self.src.device

    def to(self, device, **kw) -> "BatchedMolGraph":
        g = object.__new__(BatchedMolGraph)
        g._n, g.batch_size = self._n, self.batch_size
        g._zero_in = self._zero_in
        if self._zero_in is None:
            self.check_no_zero_in_degree_quiet()
        g._zero_in = self._zero_in
        for k in ("src", "dst", "indptr", "indices", "indptr_t", "indices_t",
                  "in_deg", "out_deg", "norm_dst", "norm_src"):
            setattr(g, k, getattr(self, k).to(device, **kw))
        g.ndata = {k: v.to(device, **kw) for k, v in self.ndata.items()}
        c = self.compact()
        g.__dict__["_compact"] = None if c is None else c.to(device, **kw)
        return g

    # ---- constructors -----------------------------------------------------------
    @classmethod
    def from_dgl(cls, g) -> "BatchedMolGraph":
        """Accept a real DGLGraph (or anything with edges()/num_nodes()/batch_size/ndata)."""
        if isinstance(g, cls):
            return g
        src, dst = g.edges()
        h = g.ndata["h"] if "h" in g.ndata else None
        return cls(src, dst, g.num_nodes(), g.batch_size, h)

    # ---- virtual-node dedup ------------------------------------------------------------
    def compact(self):
        """-> CompactMolGraph or None.  Every molecule is padded to 512 nodes with VIRTUAL nodes
        (dataset.py:214-222): isolated, one self loop, one and the same feature row -- 92 % of all
        rows.  Isolated identical rows stay identical through every GCN layer, so the layer needs
        evaluating once for them.  The compact graph keeps the real nodes plus ONE representative
        virtual node (the last row) and remembers how many rows it stands for; BatchNorm counts it
        that often (dl_batchnorm last_row_weight) and the result is expanded back to all N rows, so
        nothing is dropped or masked.  Built on the host (collate time) from CPU tensors; None when
        the graph was assembled on the device or has no such nodes."""
        c = self.__dict__.get("_compact", False)
        if c is not False:
            return c
        c = None
        h = self.ndata.get("h")
        if h is not None and not self.src.is_cuda and not h.is_cuda and self._n > 1:
            ind, outd = self.in_deg, self.out_deg
            first_src = torch.full((self._n,), -1, dtype=torch.int64)
            has = ind > 0
            first_src[has] = self.indices[self.indptr[:-1][has].long()].long()
            cand = (ind == 1) & (outd == 1) & (first_src == torch.arange(self._n))
            if int(cand.sum()) >= 2:
                rep = int(torch.nonzero(cand)[0])
                virt = cand & (h == h[rep]).all(dim=1)
                m = int(virt.sum())
                if m >= 2:
                    real_idx = torch.nonzero(~virt).flatten()
                    R = int(real_idx.numel())
                    new_id = torch.full((self._n,), -1, dtype=torch.int64)
                    new_id[real_idx] = torch.arange(R)
                    keep = ~virt[self.src]
                    src_c = torch.cat((new_id[self.src[keep]], torch.tensor([R])))
                    dst_c = torch.cat((new_id[self.dst[keep]], torch.tensor([R])))
                    assert int(src_c.min()) >= 0 and int(dst_c.min()) >= 0    # virtual nodes are isolated
                    c = CompactMolGraph(BatchedMolGraph(src_c, dst_c, R + 1, self.batch_size),
                                        torch.cat((real_idx, torch.tensor([rep]))), real_idx, m, self._n)
        self.__dict__["_compact"] = c
        return c

    # ---- device-side construction (no host sync: usable on an input pipeline's copy stream) --------
    @classmethod
    def from_edges_device(cls, src: torch.Tensor, dst: torch.Tensor, num_nodes: int, batch_size: int,
                          h: Optional[torch.Tensor] = None) -> "BatchedMolGraph":
        """Build the carrier from CUDA int64 edge lists with ``dl_csr_build`` (hand-written kernels:
        degree count, scan, stable scatter) instead of torch argsort / bincount and their host
        round-trips.  Produces exactly what ``__init__`` builds on the host."""
        if not (src.is_cuda and dst.is_cuda and src.dtype == torch.int64 and dst.dtype == torch.int64):
            raise TypeError("from_edges_device needs CUDA int64 src / dst")
        g = object.__new__(cls)
        g._n, g.batch_size = int(num_nodes), int(batch_size)
        dev, E = src.device, int(src.numel())
        g.src, g.dst = src.contiguous(), dst.contiguous()
        g.ndata = {} if h is None else {"h": h}
        i32 = dict(dtype=torch.int32, device=dev)
        g.indptr, g.indptr_t = torch.empty(g._n + 1, **i32), torch.empty(g._n + 1, **i32)
        g.indices, g.indices_t = torch.empty(E, **i32), torch.empty(E, **i32)
        g.norm_src = torch.empty(g._n, dtype=torch.float32, device=dev)
        g.norm_dst = torch.empty(g._n, dtype=torch.float32, device=dev)
        g._flags = torch.empty(2, **i32)
        g._ws = torch.empty(2 * g._n + 2 * E, **i32)
        g._zero_in = None                      # unknown until someone asks (one host read, cached)
        g.rebuild_()
        return g

    def rebuild_(self, src: Optional[torch.Tensor] = None, dst: Optional[torch.Tensor] = None) -> "BatchedMolGraph":
        """Refill this carrier's (fixed-address) CSR buffers from new edge lists of the same length:
        what a CUDA-graph-captured step needs when the next batch arrives as raw (src, dst)."""
        from . import _lib as L
        if src is not None:
            if src.numel() != self.src.numel():
                raise ValueError("rebuild_ needs edge lists of the captured length")
            self.src.copy_(src, non_blocking=True)
            self.dst.copy_(dst, non_blocking=True)
        if not hasattr(self, "_ws"):
            dev = self.src.device
            self._flags = torch.empty(2, dtype=torch.int32, device=dev)
            self._ws = torch.empty(2 * self._n + 2 * self.src.numel(), dtype=torch.int32, device=dev)
        L.call("dl_csr_build", self.src.data_ptr(), self.dst.data_ptr(), self.src.numel(), self._n,
               self.indptr.data_ptr(), self.indices.data_ptr(), self.indptr_t.data_ptr(),
               self.indices_t.data_ptr(), self.norm_src.data_ptr(), self.norm_dst.data_ptr(),
               self._flags.data_ptr(), self._ws.data_ptr())
        self._zero_in = None
        return self

    @property
    def in_deg(self):
        d = self.__dict__.get("_in_deg")
        return d if d is not None else (self.indptr[1:] - self.indptr[:-1]).long()

    @in_deg.setter
    def in_deg(self, v):
        self.__dict__["_in_deg"] = v

    @property
    def out_deg(self):
        d = self.__dict__.get("_out_deg")
        return d if d is not None else (self.indptr_t[1:] - self.indptr_t[:-1]).long()

    @out_deg.setter
    def out_deg(self, v):
        self.__dict__["_out_deg"] = v

    def check_no_zero_in_degree_quiet(self) -> None:
        if self._zero_in is None:
            f = self._flags.tolist()
            self._zero_in = bool(f[0])

    def check_no_zero_in_degree(self) -> None:
        """Mirror of the DGLError raised at ``basic_model.py:580-590``."""
        if self._zero_in is None:              # device-built: read the kernel's flags once
            f = self._flags.tolist()
            if f[1]:
                raise ValueError("edge endpoint out of range")
            self._zero_in = bool(f[0])
        if self._zero_in:
            raise Exception("There are 0-in-degree nodes in the graph, output for those nodes "
                            "will be invalid. Adding self-loop on the input graph will resolve the issue.")


class CompactMolGraph:
    """Real nodes + one representative virtual node (row R) of a BatchedMolGraph (see compact())."""

    def __init__(self, graph: BatchedMolGraph, gather: torch.Tensor, real_idx: torch.Tensor, n_virtual: int,
                 n_full: int):
        self.graph, self.gather, self.real_idx = graph, gather, real_idx
        self.n_virtual, self.n_full = int(n_virtual), int(n_full)

    def to(self, device, **kw) -> "CompactMolGraph":
        return CompactMolGraph(self.graph.to(device, **kw), self.gather.to(device, **kw),
                               self.real_idx.to(device, **kw), self.n_virtual, self.n_full)

    def tensors(self):
        g = self.graph
        return [self.gather, self.real_idx, g.indptr, g.indices, g.indptr_t, g.indices_t, g.norm_src, g.norm_dst]
