"""The dense scorer of MEVI/document_encoder.py (lines 128-132, 213-226) on libmevi_b200.

Only `compute_similarity` / `generate` are on the index hot path (they score a
query against the gathered candidate passages, main_models.py:3967-3968); the
rest of the reference class — the BERT/T5 towers — is out of scope, so
`encode_query` / `encode_passage` are not provided and `generate` requires
`p_reps`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import _lib


@dataclass
class DocEncOutput:  # same field names as the reference's ModelOutput (document_encoder.py:20-25)
    q_reps: Optional[Tensor] = None
    p_reps: Optional[Tensor] = None
    pos_p_reps: Optional[Tensor] = None
    loss: Optional[Tensor] = None
    scores: Optional[Tensor] = None


class DocumentEncoder:
    def compute_similarity(self, q_reps: Tensor, p_reps: Tensor, bmm: bool = False) -> Tensor:
        """document_encoder.py:128-132.  bmm=False: q . p^T (the hot-path form: one
        query [d] or a batch [nq,d] against passages [n,d]) in the CUDA kernel;
        bmm=True: row-wise sum(q*p) — not used by the re-rank loop, tensor ops."""
        if bmm:
            return torch.sum(q_reps * p_reps, dim=-1)
        ctx = _lib.get_context(q_reps.device.index if q_reps.is_cuda else None)
        dev = torch.device("cuda", ctx.device)
        squeeze = q_reps.dim() == 1
        q = q_reps.reshape(1, -1) if squeeze else q_reps
        q = q.to(device=dev, dtype=torch.float32).contiguous()
        p = p_reps.to(device=dev, dtype=torch.float32).contiguous()
        out = ctx.dense_scores(q, p)
        return out[0] if squeeze else out

    def generate(self, q_reps: Tensor = None, passage=None, p_reps: Tensor = None, bmm: bool = False) -> DocEncOutput:
        """document_encoder.py:213-226."""
        if p_reps is None:
            raise NotImplementedError("encode_passage (the passage tower) is out of scope: pass p_reps")
        scores = None if q_reps is None else self.compute_similarity(q_reps, p_reps, bmm)
        return DocEncOutput(scores=scores, q_reps=q_reps, p_reps=p_reps)
