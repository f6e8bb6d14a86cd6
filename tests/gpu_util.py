"""Helpers shared by the -m gpu parity tests (all calls go through the C ABI via mevi_b200._lib)."""
import numpy as np
import torch


def ctx():
    import mevi_b200

    return mevi_b200.get_context(0)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda(0)


def tensor_modes(c, d, M, K, metric="l2"):
    """Kernel modes to exercise for this shape: exact always; tensor when the library supports it."""
    modes = ["exact"]
    X = torch.zeros((8, d), device="cuda:0")
    cb = torch.zeros((M, K, d), device="cuda:0")
    try:
        c.rq_encode(X, cb, metric=metric, mode="tensor")
        modes.append("tensor")
    except Exception as e:  # MEVI_ERR_UNSUPPORTED
        if "unsupported" not in str(e).lower() and "not built" not in str(e).lower():
            raise
    return modes


def assert_topk_equivalent(s_a, i_a, s_b, i_b, rtol=1e-5, atol=1e-5, pool_scores=None):
    """Top-k lists are equal modulo score ties: same scores position by position (within tol), and
    where ids differ the two documents' scores are within tolerance of each other."""
    s_a, i_a, s_b, i_b = map(np.asarray, (s_a, i_a, s_b, i_b))
    assert s_a.shape == s_b.shape and i_a.shape == i_b.shape
    fin = np.isfinite(s_b)
    assert (np.isfinite(s_a) == fin).all(), "padding differs"
    np.testing.assert_allclose(s_a[fin], s_b[fin], rtol=rtol, atol=atol)
    assert ((i_a < 0) == (i_b < 0)).all()
    n_tie = 0
    for q in range(i_a.shape[0]):
        if (i_a[q] == i_b[q]).all():
            continue
        set_a, set_b = set(i_a[q][i_a[q] >= 0].tolist()), set(i_b[q][i_b[q] >= 0].tolist())
        only = (set_a - set_b) | (set_b - set_a)
        # boundary ties: documents present in only one list must score within tol of the k-th score
        kth = s_b[q][fin[q]][-1] if fin[q].any() else 0.0
        for doc in only:
            assert pool_scores is not None, f"query {q}: id sets differ ({sorted(only)[:6]}...)"
            sc = pool_scores(q, doc)
            assert abs(sc - kth) <= atol + rtol * abs(kth), f"query {q}: doc {doc} score {sc} vs k-th {kth}"
        # order swaps inside the list must be between near-equal scores
        pos_b = {int(d_): p for p, d_ in enumerate(i_b[q])}
        for p, d_ in enumerate(i_a[q]):
            if d_ >= 0 and int(d_) in pos_b and pos_b[int(d_)] != p:
                assert abs(s_b[q][pos_b[int(d_)]] - s_a[q][p]) <= atol + rtol * abs(s_a[q][p])
                n_tie += 1
    return n_tie
