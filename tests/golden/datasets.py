"""Seeded synthetic datasets shared by the golden-vector generator and the tests.

numpy's legacy `RandomState` is used on purpose: its streams are frozen across
numpy versions, so the data regenerate bit-identically on the GPU box (each
case's sha256 is pinned in its meta.json and re-checked by the tests).
"""
from __future__ import annotations

import hashlib

import numpy as np

CASES = {
    # name: (kind, N, d, M, bits, seed)
    "gauss768": ("gauss", 3000, 768, 4, 5, 1234),
    "mix768": ("mix", 3000, 768, 4, 5, 99),
    "small64": ("gauss", 2000, 64, 3, 4, 7),
    # BASELINE.json configs[0]: synthetic 100k x 768 (make_golden_100k.py; codebook + codes only)
    "gauss100k": ("gauss", 100000, 768, 4, 5, 1234),
}
QUERY_SEED = 4321
N_QUERIES = 32


def make_docs(kind: str, n: int, d: int, seed: int) -> np.ndarray:
    rs = np.random.RandomState(seed)
    if kind == "gauss":
        return rs.standard_normal((n, d)).astype(np.float32)
    if kind == "mix":
        centers = rs.standard_normal((64, d)).astype(np.float32)
        lab = rs.randint(0, 64, size=n)
        noise = rs.standard_normal((n, d)).astype(np.float32)
        return (centers[lab] + np.float32(0.3) * noise).astype(np.float32)
    raise ValueError(kind)


def make_queries(d: int, n: int = N_QUERIES, seed: int = QUERY_SEED) -> np.ndarray:
    return np.random.RandomState(seed).standard_normal((n, d)).astype(np.float32)


def case_docs(name: str) -> np.ndarray:
    kind, n, d, _, _, seed = CASES[name]
    return make_docs(kind, n, d, seed)


def sha256(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
