"""Import the UNMODIFIED reference `MEVI/pq.py` in the authoring container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  `/root/reference` exists only in
the authoring container, never on the GPU box, so this module is used solely by
`tests/golden/make_golden.py` (to mint fixtures) and by the optional
`tests/test_oracle_vs_reference.py` (skipped when the checkout is absent).

`MEVI/pq.py:6` does `import faiss` at module scope; faiss is not installed in
this image and none of the RQ k-means / encode / beam-search functions touch it
(only `build_faiss_index` / `codebook_from_index`, pq.py:143-198), so an empty
stub module is registered before the import.  Nothing else is patched.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MEVI_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "MEVI", "pq.py"))


def load_reference_pq():
    """Return the reference's `pq` module object (cached in sys.modules)."""
    name = "_mevi_reference_pq"
    if name in sys.modules:
        return sys.modules[name]
    if not reference_available():
        raise FileNotFoundError(f"reference checkout not found under {REFERENCE_ROOT}")
    if "faiss" not in sys.modules:
        try:
            import faiss  # noqa: F401
        except Exception:
            sys.modules["faiss"] = types.ModuleType("faiss")
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, "MEVI", "pq.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_gen_sampled():
    """`dataprocess/msmarco_passage/gen_sampled_to_full.py` — torch-only second
    copy of the greedy encode (lines 20-22, 65-86)."""
    name = "_mevi_reference_gen_sampled"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(REFERENCE_ROOT, "dataprocess", "msmarco_passage", "gen_sampled_to_full.py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
