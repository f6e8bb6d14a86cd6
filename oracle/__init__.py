"""CPU oracle for the MEVI index hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mevi_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may.  It is the checker, never the product.
"""
