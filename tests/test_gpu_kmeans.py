"""k-means step parity: assignment, per-centroid sums/counts (fused buffer), update, residual."""
import numpy as np
import pytest
import torch

from oracle import oracle
from gpu_util import ctx, dev

pytestmark = pytest.mark.gpu


def _step(c, X, C, mode="exact"):
    K, d = C.shape
    buf = torch.empty(K * d + K, device="cuda:0")
    assign = torch.empty(len(X), dtype=torch.int32, device="cuda:0")
    inertia = torch.zeros(1, dtype=torch.float64, device="cuda:0")
    c.kmeans_step(dev(X), dev(C), buf, assign=assign, inertia=inertia, mode=mode)
    b = buf.cpu().numpy()
    return assign.cpu().numpy(), b[: K * d].reshape(K, d), b[K * d :], float(inertia.item())


@pytest.mark.parametrize("mode", ["exact", "auto"])
@pytest.mark.parametrize("shape", [(5000, 768, 32), (3000, 64, 16), (4097, 256, 64), (2000, 128, 100), (30000, 768, 32)])
def test_one_lloyd_step_vs_float64_oracle(shape, mode):
    """`auto` = what the trainer runs: the tensor assignment kernel (K1 at M = 1) where the shape allows it (the 768- and
    256-wide cases with n >= 4096 here), the direct fp32 kernel otherwise."""
    n, d, K = shape
    rs = np.random.RandomState(5)
    X = rs.standard_normal((n, d)).astype(np.float32)
    C = X[rs.choice(n, K, replace=False)].copy() * 0.5
    c = ctx()
    assign, sums, counts, inertia = _step(c, X, C, mode=mode)
    a_ref, s_ref, c_ref, i_ref, gaps = oracle.lloyd_step(X, C)
    diff = np.nonzero(assign != a_ref)[0]
    assert (gaps[diff] <= oracle.TIE_EPS_DEFAULT * 4).all(), "assignment differs at a non-tie"
    if len(diff) == 0:
        assert np.array_equal(counts, c_ref)
        np.testing.assert_allclose(sums, s_ref, rtol=2e-5, atol=2e-4 if n < 10000 else 2e-3)
    assert abs(inertia - i_ref) <= (1e-5 if mode == "exact" else 5e-5) * i_ref  # tensor path: fp16-split distances
    assert counts.sum() == n


def test_accumulation_is_deterministic():
    rs = np.random.RandomState(6)
    X = rs.standard_normal((20000, 768)).astype(np.float32)
    C = X[:32].copy()
    c = ctx()
    a1, s1, c1, _ = _step(c, X, C)
    a2, s2, c2, _ = _step(c, X, C)
    assert np.array_equal(a1, a2) and np.array_equal(s1, s2) and np.array_equal(c1, c2)


def test_update_and_residual():
    rs = np.random.RandomState(7)
    n, d, K = 3000, 768, 32
    X = rs.standard_normal((n, d)).astype(np.float32)
    C = X[:K].copy()
    C[5] = 100.0  # far away: ends up empty
    c = ctx()
    assign, sums, counts, _ = _step(c, X, C)
    assert counts[5] == 0
    Cd = dev(C)
    buf = dev(np.concatenate([sums.ravel(), counts]).astype(np.float32))
    n_empty = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    c.kmeans_update(buf, Cd, n_empty)
    newC = Cd.cpu().numpy()
    assert int(n_empty.item()) == 1 and np.array_equal(newC[5], C[5])  # empty cluster keeps its centroid
    nz = counts > 0
    np.testing.assert_allclose(newC[nz], sums[nz] / counts[nz, None], rtol=1e-6)
    R = dev(X)
    c.residual_update(R, Cd, dev(assign))
    assert np.array_equal(R.cpu().numpy(), X - newC[assign])  # fp32 elementwise, pq.py:591-593


def test_strided_assign_column():
    rs = np.random.RandomState(8)
    X = rs.standard_normal((1000, 64)).astype(np.float32)
    C = X[:16].copy()
    c = ctx()
    codes = torch.full((1000, 4), -7, dtype=torch.int32, device="cuda:0")
    buf = torch.empty(16 * 64 + 16, device="cuda:0")
    c.kmeans_step(dev(X), dev(C), buf, assign=codes[:, 2], assign_stride=4, mode="exact")
    out = codes.cpu().numpy()
    assert (out[:, [0, 1, 3]] == -7).all()
    assert np.array_equal(out[:, 2], oracle.lloyd_step(X, C)[0])


def test_trainer_quality_vs_reference_codebook(case):
    """Full-batch Lloyd on the golden data reaches a quantisation MSE no worse than the codebook the
    reference trained with sklearn MiniBatchKMeans (parity for training = MSE, not bits)."""
    from mevi_b200.pq import ProductQuantization

    pq = ProductQuantization("rq", case.M, case.meta["bits"], "l2", case.d, "kmeans", "grad")
    pq.kernel_mode = "exact"
    pq.initialize(None, case.X, 0, 41, None, 1024)
    assert pq.get_preds and pq.last_preds.shape == (case.n, case.M) and pq.last_preds.dtype == np.int32
    cb = pq.codebook.detach().numpy()
    mse_new = oracle.quantisation_mse(case.X, cb, pq.last_preds)
    mse_ref = oracle.quantisation_mse(case.X, case.codebook, case.codes)
    print(case.name, "mse new", mse_new, "ref", mse_ref)
    assert mse_new <= mse_ref * 1.02
    # last_preds are the greedy codes of the final codebook (modulo ties), like fit_predict in the reference
    rep = oracle.classify_code_mismatches(case.X, cb, oracle.rq_encode(case.X, cb), pq.last_preds)
    assert rep["n_hard"] == 0
    clus, mapping = pq.get_document_cluster_simple(True)
    assert len(mapping) == case.n and sum(len(v) for v in clus.values()) == case.n


@pytest.mark.parametrize("n,d,K", [(20000, 768, 32), (4097, 768, 32), (70001, 256, 32), (9000, 512, 32)])
def test_fused_step_assigns_and_accumulates_in_one_pass(n, d, K):
    """mevi_kmeans_step_fused: new assignment under C + per-centroid sums|counts of the rows under the PREVIOUS
    assignment, in one read of the shard.  Against the two-pass kernels (assignment bit-equal, sums to fp32 rounding),
    a float64 oracle, and itself (bit-reproducible)."""
    rs = np.random.RandomState(n)
    X = rs.standard_normal((n, d)).astype(np.float32)
    X[: n // 3] += 2.0 * rs.standard_normal((1, d)).astype(np.float32)  # skew: a third of the rows share a cluster
    C0 = X[rs.choice(n, K, replace=False)].copy()
    C1 = (C0 + 0.05 * rs.standard_normal((K, d))).astype(np.float32)
    c = ctx()
    Xd, C0d, C1d = dev(X), dev(C0), dev(C1)
    buf = torch.empty(K * d + K, device="cuda:0")
    prev = torch.empty(n, dtype=torch.int32, device="cuda:0")
    c.kmeans_step(Xd, C0d, buf, assign=prev, mode="tensor")
    want_assign = torch.empty(n, dtype=torch.int32, device="cuda:0")
    c.kmeans_step(Xd, C1d, buf, assign=want_assign, mode="tensor")
    want_prev_sums = c.accumulate_by_code(Xd, prev, K).clone()
    cur = torch.empty(n, dtype=torch.int32, device="cuda:0")
    got = torch.empty(K * d + K, device="cuda:0")
    inertia = torch.zeros(1, dtype=torch.float64, device="cuda:0")
    c.kmeans_step_fused(Xd, C1d, prev, cur, got, inertia=inertia)
    c.check()
    assert torch.equal(cur, want_assign)
    assert torch.equal(got[K * d :], want_prev_sums[K * d :])  # counts
    np.testing.assert_allclose(got[: K * d].cpu().numpy(), want_prev_sums[: K * d].cpu().numpy(), rtol=2e-5, atol=2e-3)
    pa = prev.cpu().numpy()
    s64 = np.zeros((K, d), np.float64)
    np.add.at(s64, pa, X.astype(np.float64))
    np.testing.assert_allclose(got[: K * d].view(K, d).cpu().numpy(), s64, rtol=2e-5, atol=2e-3)
    _, _, _, i_ref, _ = oracle.lloyd_step(X, C1)
    assert abs(float(inertia.item()) - i_ref) <= 1e-5 * i_ref
    got2 = torch.empty_like(got)
    cur2 = torch.empty_like(cur)
    c.kmeans_step_fused(Xd, C1d, prev, cur2, got2)
    assert torch.equal(got, got2) and torch.equal(cur, cur2)


def test_trainer_one_pass_iterations_match_two_pass_training(monkeypatch):
    """train_rq_lloyd with one-pass iterations (fused assign + accumulate, changed rows moved afterwards) against the
    same training with two passes per iteration: same codes, same codebook up to fp32 summation order."""
    from mevi_b200 import trainer

    rs = np.random.RandomState(21)
    centers = rs.standard_normal((40, 768)).astype(np.float32)
    X = (centers[rs.randint(0, 40, 30000)] + 0.7 * rs.standard_normal((30000, 768))).astype(np.float32)
    out = {}
    for how in ("delta", "fused", "twopass"):
        monkeypatch.setattr(trainer, "LLOYD_ITERATION", how)
        cb, codes = trainer.train_rq_lloyd(X, M=2, K=32, seed=41, iters=8, tol=None, device_index=0)
        info = trainer.train_rq_lloyd.last_info
        out[how] = (cb.cpu().numpy(), codes, info)
    lv = lambda how: out[how][2]["levels"][0]
    assert lv("delta")["delta_iters"] == 7 and lv("fused")["fused_iters"] == 7 and lv("twopass")["two_pass_iters"] == 8
    assert lv("delta")["fused_iters"] == 0 and lv("twopass")["delta_iters"] == 0
    assert lv("delta")["changed_rows"] > 0
    assert abs(lv("delta")["changed_rows"] - lv("fused")["changed_rows"]) <= 0.01 * lv("fused")["changed_rows"] + 5
    m0 = oracle.quantisation_mse(X, out["twopass"][0], out["twopass"][1])
    for how in ("delta", "fused"):
        np.testing.assert_allclose(out[how][0], out["twopass"][0], rtol=1e-4, atol=1e-4)
        assert (out[how][1] != out["twopass"][1]).any(1).mean() < 2e-3
        m1 = oracle.quantisation_mse(X, out[how][0], out[how][1])
        assert abs(m1 - m0) <= 1e-4 * m0


@pytest.mark.parametrize("n,d,K,mode", [(20000, 768, 32, "auto"), (4097, 768, 32, "exact"), (70001, 256, 32, "auto"),
                                         (9000, 512, 16, "auto"), (3000, 64, 8, "auto"), (6000, 32, 256, "auto"),
                                         (5000, 24, 64, "auto")])
def test_delta_step_corrects_running_sums_by_the_moved_rows(n, d, K, mode):
    """mevi_kmeans_step_delta: one assignment pass + float64 running sums|counts corrected by the rows whose assignment
    changed.  Five chained iterations against a float64 oracle of the sums under the NEW assignment, the two-pass step
    (assignment bit-equal), and itself (bit-reproducible)."""
    rs = np.random.RandomState(n + 1)
    X = rs.standard_normal((n, d)).astype(np.float32)
    X[: n // 3] += 2.0 * rs.standard_normal((1, d)).astype(np.float32)
    C = X[rs.choice(n, K, replace=False)].copy()
    c = ctx()
    Xd = dev(X)

    def run():
        Cd = dev(C)
        buf = torch.empty(K * d + K, device="cuda:0")
        a, b = (torch.empty(n, dtype=torch.int32, device="cuda:0") for _ in range(2))
        master = torch.empty(K * d + K, dtype=torch.float64, device="cuda:0")
        nchg = torch.zeros(1, dtype=torch.int32, device="cuda:0")
        inertia = torch.zeros(1, dtype=torch.float64, device="cuda:0")
        c.kmeans_step(Xd, Cd, buf, assign=a, mode=mode)
        master.copy_(buf)
        c.kmeans_update(buf, Cd)
        trace = []
        for it in range(5):
            c.kmeans_step_delta(Xd, Cd, a, b, master, buf, n_changed=nchg, inertia=inertia, mode=mode)
            trace.append((b.clone(), buf.clone(), int(nchg.item()), float(inertia.item()), Cd.clone(), a.clone()))
            c.kmeans_update(buf, Cd)
            a, b = b, a
        c.check()
        return trace

    t1, t2 = run(), run()
    moved_total = 0
    for (cur, buf, nchg, inertia, Cd, prev), (cur2, buf2, nchg2, _, _, _) in zip(t1, t2):
        assert torch.equal(cur, cur2) and nchg == nchg2
        if K <= 64:  # bit-reproducible (for K > 64 the FIRST, two-pass step accumulates with float atomics: kmeans.cu)
            assert torch.equal(buf, buf2)
        else:
            assert torch.allclose(buf, buf2, rtol=1e-5, atol=1e-3)
        want = torch.empty(n, dtype=torch.int32, device="cuda:0")
        tmp = torch.empty(K * d + K, device="cuda:0")
        c.kmeans_step(Xd, Cd, tmp, assign=want, mode=mode)
        assert torch.equal(cur, want)
        assert nchg == int((cur != prev).sum().item())
        moved_total += nchg
        assert torch.equal(buf[K * d :], tmp[K * d :])  # counts are exact
        s64 = np.zeros((K, d), np.float64)
        np.add.at(s64, cur.cpu().numpy(), X.astype(np.float64))
        np.testing.assert_allclose(buf[: K * d].view(K, d).cpu().numpy(), s64, rtol=2e-5, atol=2e-3)
        _, _, _, i_ref, _ = oracle.lloyd_step(X, Cd.cpu().numpy())
        assert abs(inertia - i_ref) <= 5e-5 * i_ref  # a convergence statistic: the tensor path's fp16-split distances
    assert moved_total > 0
