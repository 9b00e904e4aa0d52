"""Packed wire format for the LLM embeddings.

The reference's ``multimodality_collate_func`` (``utils.py:326-334``) pads each sample's embedding
rows on the HOST -- ``tail_pad`` for the drug (``:304-312``), ``repeat_pad`` for the protein, which
tiles the (L+2, 640) block until 2304 rows are filled (``:314-324``) -- and ships the dense fp32
tensors: 6.9 MB per pair, which caps any GPU step at ~7.7 k pairs/s over PCIe.  The rows themselves
are ~1.5 MB per pair.  Here the host concatenates the untiled rows into one pinned buffer
(:class:`PackedRows`) and the padding / tiling runs on the device (``dl_expand_rows``), producing
bit-identical dense tensors for the model.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import torch

from . import kernels as K


@dataclass
class PackedRows:
    """Row blocks of one padded (B, maxsize, C) collate tensor, stored back to back."""
    rows: torch.Tensor        # [sum_b R_b, C] fp32
    offsets: torch.Tensor     # [B + 1] int32, offsets[b+1] - offsets[b] = R_b
    maxsize: int
    repeat: bool              # True: utils.repeat_pad, False: utils.tail_pad

    @property
    def batch(self) -> int:
        return self.offsets.numel() - 1

    def nbytes(self) -> int:
        return self.rows.numel() * self.rows.element_size() + self.offsets.numel() * self.offsets.element_size()

    def to(self, device, non_blocking: bool = False) -> "PackedRows":
        return PackedRows(self.rows.to(device, non_blocking=non_blocking),
                          self.offsets.to(device, non_blocking=non_blocking), self.maxsize, self.repeat)

    def dense(self, out: torch.Tensor = None) -> torch.Tensor:
        """The reference collate's dense tensor, built on the device (rows/offsets must be CUDA)."""
        if out is None:
            out = torch.empty((self.batch, self.maxsize, self.rows.shape[1]), dtype=torch.float32,
                              device=self.rows.device)
        return K.expand_rows(self.rows, self.offsets, out, self.repeat)


def pack_rows(blocks: Sequence[torch.Tensor], maxsize: int, repeat: bool, pin: bool = True) -> PackedRows:
    """Host side: concatenate the per-sample (R_b, C) blocks (what ``l['prot'].x`` / ``l['drug'].x``
    are in the reference collate) into one (optionally pinned) buffer."""
    if len(blocks) == 0:
        raise ValueError("pack_rows: empty batch")
    C = blocks[0].shape[-1]
    counts = []
    for a in blocks:
        if a.dim() != 2 or a.shape[1] != C:
            raise ValueError("pack_rows: every block must be (rows, C) with one feature width")
        if not repeat and a.shape[0] > maxsize:
            raise ValueError(f"tail_pad: a block of {a.shape[0]} rows does not fit maxsize {maxsize} "
                             "(the reference raises here too)")
        counts.append(int(a.shape[0]))
    total = sum(counts)
    pin = pin and torch.cuda.is_available()
    rows = torch.empty((total, C), dtype=torch.float32, pin_memory=pin)
    offsets = torch.zeros(len(blocks) + 1, dtype=torch.int32, pin_memory=pin)
    o = 0
    for i, a in enumerate(blocks):
        rows[o:o + counts[i]].copy_(a)
        o += counts[i]
        offsets[i + 1] = o
    return PackedRows(rows, offsets, int(maxsize), bool(repeat))


def pack_llm(llm: Sequence[dict], drug_max: int = 512, prot_max: int = 9 * 256, pin: bool = True):
    """The LLM part of ``multimodality_collate_func``: ``llm`` is the per-sample list of
    ``{'drug': obj with .x, 'prot': obj with .x}`` the reference dataset yields (tensors are accepted
    in place of the ``.x`` holders).  Returns (drug PackedRows, protein PackedRows)."""
    def x_of(v):
        return v.x if hasattr(v, "x") else v
    d = pack_rows([x_of(l["drug"]) for l in llm], drug_max, repeat=False, pin=pin)
    p = pack_rows([x_of(l["prot"]) for l in llm], prot_max, repeat=True, pin=pin)
    return d, p
