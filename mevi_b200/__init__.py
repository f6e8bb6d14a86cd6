"""mevi_b200 — B200 (sm_100a) implementation of MEVI's index hot path.

Drop-in Python mirrors of the reference's entry points for this path
(`pq.ProductQuantization`, `faiss_search.read/search/to_file`,
`document_encoder.DocumentEncoder.compute_similarity/generate`, the
cluster-restricted re-rank of `main_models.py:3911-4053`) on top of
`libmevi_b200.so`, a C-ABI library of hand-written CUDA kernels
(include/mevi_b200.h).  There is no CPU fallback: every compute call raises if
the library or a CUDA device is missing.
"""
from ._lib import Context, MeviError, get_context, library_path, load_library  # noqa: F401

__all__ = ["Context", "MeviError", "get_context", "library_path", "load_library"]
__version__ = "0.1.0"
