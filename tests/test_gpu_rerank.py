"""Inverted lists, cluster-restricted re-rank, top-k merge and dense scorer parity."""
import numpy as np
import pytest
import torch

from oracle import oracle
from gpu_util import assert_topk_equivalent, ctx, dev

pytestmark = pytest.mark.gpu


def test_inverted_lists_equal_reference_dicts(case):
    from mevi_b200.rerank import ClusterIndex

    idx = ClusterIndex.from_codes(case.codes, case.K)
    clus, mapping = idx.to_dicts()
    ref_clus, ref_map = case.pickle("rqclus.pkl"), case.pickle("rqmapping.pkl")
    assert clus == ref_clus and mapping == ref_map
    # ... in the reference's insertion order too, so the pickles written from the device CSR are the reference's bytes
    import pickle

    assert list(clus) == list(ref_clus) and list(mapping) == list(ref_map)
    assert pickle.dumps(clus) == pickle.dumps(ref_clus) and pickle.dumps(mapping) == pickle.dumps(ref_map)
    # a shard with an id base: global doc ids, same rule
    half = case.n // 2
    c2, m2 = ClusterIndex.from_codes(case.codes[half:], case.K, id_base=half).to_dicts()
    assert min(m2) == half and all(m2[i] == ref_map[i] for i in m2) and sum(len(v) for v in c2.values()) == case.n - half
    idx2 = ClusterIndex.from_cluster_dict(ref_clus, case.K)
    assert torch.equal(idx2.leaf_keys, idx.leaf_keys) and torch.equal(idx2.leaf_offsets, idx.leaf_offsets)
    assert torch.equal(idx2.leaf_docids, idx.leaf_docids)


@pytest.mark.parametrize("leaf_ordered", [True, False])
@pytest.mark.parametrize("nb,k", [(10, 100), (100, 100), (10, 1000), (100, 7)])
def test_rerank_matches_oracle(case, nb, k, leaf_ordered):
    from mevi_b200.rerank import ClusterIndex, ClusterReranker

    dec = case.load(f"beam{nb}_labels.npy")
    clus = case.pickle("rqclus.pkl")
    rr = ClusterReranker(dev(case.X), ClusterIndex.from_codes(case.codes, case.K), leaf_ordered=leaf_ordered)
    scores, ids, ncand = rr.rerank(case.Q, dec, topk=k)
    scores, ids, ncand = scores.cpu().numpy(), ids.cpu().numpy(), ncand.cpu().numpy()
    ref = oracle.cluster_rerank(case.Q, case.X, clus, dec, topk=k)
    s_ref = np.full((len(ref), k), -np.inf, np.float32)
    i_ref = np.full((len(ref), k), -1, np.int64)
    for q, (d_, s_, nd) in enumerate(ref):
        s_ref[q, : len(s_)] = s_
        i_ref[q, : len(d_)] = d_
        assert ncand[q] == nd
    X64, Q64 = case.X.astype(np.float64), case.Q.astype(np.float64)
    assert_topk_equivalent(scores, ids, s_ref, i_ref, rtol=1e-5, atol=1e-4,
                           pool_scores=lambda q, doc: float(X64[doc] @ Q64[q]))
    assert (np.diff(np.where(np.isfinite(scores), scores, -1e30), axis=1) <= 0).all()


@pytest.mark.parametrize("plan", ["device", "tiles", "prefix"])
@pytest.mark.parametrize("nb,k,boot_min", [(10, 100, 128), (100, 100, 128), (100, 7, 128), (100, 100, 100000)])
def test_grouped_tensor_rerank_matches_oracle(case, nb, k, boot_min, plan):
    """K3g: every leaf read once and scored against all the queries that chose it (tcgen05 prefilter + exact fp32
    re-score) must return what the per-query loop of main_models.py:3915-4014 returns."""
    from mevi_b200.rerank import ClusterIndex, ClusterReranker

    if case.d % 64:
        pytest.skip("tensor path needs d % 64 == 0")
    dec = case.load(f"beam{nb}_labels.npy")
    clus = case.pickle("rqclus.pkl")
    rr = ClusterReranker(dev(case.X), ClusterIndex.from_codes(case.codes, case.K), mode="grouped")
    # small corpus: still exercise the threshold-free bootstrap round + three more rounds; boot_min = 100000 makes every
    # query "weak" (first thresholds from the streaming kernel's exact prefix top-k instead)
    rr.BOOTSTRAP_ROWS, rr.ROUND_ROWS, rr.BOOTSTRAP_MIN = (256 if boot_min < 1000 else 32), (600, 3000), boot_min
    rr.PASS_BUDGET = 16  # (a 3,000-document corpus: without this no query would ever need the streaming bootstrap)
    # "tiles" (the default plan): bootstrap = first tile of the 3 leading leaves, everything else in the second round
    rr.PLAN, rr.BOOT_LEAVES = plan, 3
    scores, ids, ncand = rr.rerank(case.Q, dec, topk=k)
    assert rr.last_path == "grouped" and rr.last_failed_queries == 0
    if case.name == "gauss768" and plan == "prefix":  # (the other cases' leaves are so small that 32 rows already hold all candidates)
        assert (rr.last_weak_queries > 0) == (boot_min > 1000)
    scores, ids, ncand = scores.cpu().numpy(), ids.cpu().numpy(), ncand.cpu().numpy()
    ref = oracle.cluster_rerank(case.Q, case.X, clus, dec, topk=k)
    s_ref = np.full((len(ref), k), -np.inf, np.float32)
    i_ref = np.full((len(ref), k), -1, np.int64)
    for q, (d_, s_, nd) in enumerate(ref):
        s_ref[q, : len(s_)] = s_
        i_ref[q, : len(d_)] = d_
        assert ncand[q] == nd
    X64, Q64 = case.X.astype(np.float64), case.Q.astype(np.float64)
    assert_topk_equivalent(scores, ids, s_ref, i_ref, rtol=1e-5, atol=1e-4,
                           pool_scores=lambda q, doc: float(X64[doc] @ Q64[q]))


@pytest.mark.parametrize("maxg_sample,maxg_last", [(1, 1), (1, 4), (2, 2), (4, 4)])
def test_device_plan_covers_every_pair_tile_exactly_once(maxg_sample, maxg_last):
    """mevi_rerank_grouped_plan (csrc/rerank_plan.cu): every (query, leaf) pair meets every tile of its leaf exactly once,
    in the round its leaf rank assigns; candidate totals and weak-sample flags equal the torch arithmetic."""
    from mevi_b200.rerank import GROUP_COLS, build_leaf_tiles

    c = ctx()
    rs = np.random.RandomState(1)
    n_leaves = 300
    sizes = rs.randint(1, 700, size=n_leaves)
    sizes[2], sizes[5], sizes[6], sizes[7] = 1, 128, 129, 5000
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)).cuda()
    _, _, lt0, _ = build_leaf_tiles(off)
    nq, L, boot = 700, 12, (2, 5)
    ql_h = np.stack([rs.choice(40, size=L, replace=False) for _ in range(nq)]).astype(np.int32)  # 40 hot leaves: full groups
    ql_h[::3, 7:] = rs.randint(40, n_leaves, size=ql_h[::3, 7:].shape)
    ql_h[4, 2] = -1
    ql_h[9, :] = -1
    ql_h[11, :] = 2  # one tiny leaf asked for twelve times (few candidates: needs no sample)
    ql = torch.from_numpy(ql_h).cuda()
    ncand, weak, rounds, n_weak = c.rerank_grouped_plan(ql, off, lt0, boot, 900, 500, maxg_sample, maxg_last, pass_budget=6144)
    want_ncand = np.where(ql_h >= 0, sizes[np.clip(ql_h, 0, None)], 0).sum(1)
    assert np.array_equal(ncand.cpu().numpy(), want_ncand)
    boot_rows = np.where(ql_h[:, :5] >= 0, np.minimum(sizes[np.clip(ql_h[:, :5], 0, None)], 128), 0).sum(1)
    limit = np.minimum(900, np.maximum(2 * 500, want_ncand * 500 // 6144))  # boot_min_rows 900, k 500
    want_weak = (want_ncand > 6144) & (boot_rows < limit) & (want_ncand > boot_rows)
    assert np.array_equal(weak.cpu().numpy().astype(bool), want_weak) and n_weak == int(want_weak.sum()) and n_weak >= 1
    assert len(rounds) == 3
    lt0_h = lt0.cpu().numpy()
    seen = []
    for r, (n_items, n_groups) in enumerate(rounds):
        assert n_items > 0 and n_groups > 0
        it, ig, gq = c.rerank_grouped_plan_fill(r, n_items, n_groups, ql.device)
        it, ig, gq = it.cpu().numpy(), ig.cpu().numpy(), gq.cpu().numpy().reshape(-1, GROUP_COLS)
        cap = maxg_last if r == 2 else maxg_sample
        for i in range(n_items):
            t, g0, ng = int(it[i]), int(ig[i]) & 0xFFFFFF, int(ig[i]) >> 24
            assert 1 <= ng <= cap and g0 + ng <= n_groups
            leaf = int(np.searchsorted(lt0_h, t, side="right")) - 1
            for g in range(g0, g0 + ng):
                seen += [(r, int(q), leaf, t) for q in gq[g][gq[g] >= 0]]
    # query 11 lists leaf 2 twelve times: the reference scores it twelve times too (a list, not a set)
    from collections import Counter

    want = Counter()
    for q in range(nq):
        for j in range(L):
            leaf = int(ql_h[q, j])
            if leaf < 0:
                continue
            t0, t1 = int(lt0_h[leaf]), int(lt0_h[leaf + 1])
            want[(0 if j < 2 else (1 if j < 5 else 2), q, leaf, t0)] += 1
            for t in range(t0 + 1, t1):
                want[(2, q, leaf, t)] += 1
    assert Counter(seen) == want


def test_grouped_rerank_falls_back_when_the_margin_window_overflows(gauss):
    """Near-duplicate documents put more candidates inside the fp16 margin than the buffers keep: the grouped path must
    notice and the call must still return the exact answer (through the streaming kernel)."""
    from mevi_b200.rerank import ClusterIndex, ClusterReranker

    X = gauss.X.copy()
    nd = 2500  # near-duplicates: all of them inside the fp16 margin window, far more than the 512 the compaction keeps
    X[:nd] = X[0] + 1e-6 * np.arange(nd, dtype=np.float32)[:, None]
    codes = np.ascontiguousarray(gauss.codes.copy())
    codes[:nd] = codes[0]
    clus, _ = oracle.document_cluster(codes)
    dec = np.repeat(codes[None, :1, :], 4, axis=0)  # four queries, all asking for that one leaf
    Q = np.repeat(X[:1], 4, axis=0) * np.float32(1.0)
    rr = ClusterReranker(dev(X), ClusterIndex.from_codes(codes, gauss.K), mode="grouped")
    scores, ids, ncand = rr.rerank(Q, dec, topk=100)
    assert rr.last_path == "stream"
    ref = oracle.cluster_rerank(Q, X, clus, dec, topk=100)
    for q, (d_, s_, nd) in enumerate(ref):
        assert int(ncand[q]) == nd
        np.testing.assert_allclose(scores[q].cpu().numpy(), s_, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("plan", ["device", "tiles", "prefix"])
def test_grouped_rerank_reruns_only_the_queries_whose_guarantee_failed(gauss, plan):
    """A few queries hit a leaf of near-duplicate documents (margin window overflow), the others do not: only those few
    go through the streaming kernel, and every query's answer is the exact one."""
    from mevi_b200.rerank import ClusterIndex, ClusterReranker

    X = gauss.X.copy()
    nd = 2500
    X[:nd] = X[0] + 1e-6 * np.arange(nd, dtype=np.float32)[:, None]
    codes = np.ascontiguousarray(gauss.codes.copy())
    codes[:nd] = codes[0]
    clus, _ = oracle.document_cluster(codes)
    dec = gauss.load("beam10_labels.npy").copy()          # 32 ordinary queries ...
    Q = gauss.Q.copy()
    dec[:3, 0, :] = codes[0]                              # ... three of which also ask for the duplicate leaf
    Q[:3] = X[0]
    rr = ClusterReranker(dev(X), ClusterIndex.from_codes(codes, gauss.K), mode="grouped")
    rr.BOOTSTRAP_ROWS, rr.ROUND_ROWS, rr.BOOTSTRAP_MIN = 256, (600,), 128
    rr.PLAN, rr.BOOT_LEAVES = plan, 2
    scores, ids, ncand = rr.rerank(Q, dec, topk=100)
    assert rr.last_path == "grouped+stream" and 1 <= rr.last_failed_queries <= 3
    ref = oracle.cluster_rerank(Q, X, clus, dec, topk=100)
    for q, (d_, s_, nd_) in enumerate(ref):
        assert int(ncand[q]) == nd_
        np.testing.assert_allclose(scores[q, : len(s_)].cpu().numpy(), s_, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("leaf_ordered", [True, False])
def test_rerank_split_path_and_empty_leaves(gauss, leaf_ordered):
    """Few queries -> several CTAs per query + merge; leaves that hold no document are legal."""
    from mevi_b200.rerank import ClusterIndex, ClusterReranker

    dec = gauss.load("beam100_labels.npy")[:2].copy()
    dec[0, :50] = 31  # (31,31,31,31) almost surely empty
    clus = gauss.pickle("rqclus.pkl")
    rr = ClusterReranker(dev(gauss.X), ClusterIndex.from_codes(gauss.codes, gauss.K), leaf_ordered=leaf_ordered)
    scores, ids, ncand = rr.rerank(gauss.Q[:2], dec, topk=50)
    ref = oracle.cluster_rerank(gauss.Q[:2], gauss.X, clus, dec, topk=50)
    for q, (d_, s_, nd) in enumerate(ref):
        assert int(ncand[q]) == nd
        kk = len(d_)
        np.testing.assert_allclose(scores[q, :kk].cpu().numpy(), s_, rtol=1e-5, atol=1e-4)
        assert (ids[q, kk:] == -1).all()
    # a query whose leaves are all empty
    dec2 = np.full((1, 10, gauss.M), 31, dtype=np.int64)
    s2, i2, n2 = rr.rerank(gauss.Q[:1], dec2, topk=10)
    assert int(n2[0]) == 0 and (i2 == -1).all() and torch.isinf(s2).all()


def test_topk_merge_matches_numpy():
    rs = np.random.RandomState(2)
    S, nq, k = 8, 37, 100
    s = rs.standard_normal((S, nq, k)).astype(np.float32)
    s[:, :, 10:12] = 0.5  # exact score ties across shards -> id order decides
    i = rs.permutation(S * nq * k).reshape(S, nq, k).astype(np.int64)
    s[3, :, 90:] = -np.inf
    i[3, :, 90:] = -1
    ms, mi = ctx().topk_merge(dev(s), dev(i))
    for q in range(nq):
        fs, fi = s[:, q].ravel(), i[:, q].ravel()
        keep = fi >= 0
        order = np.lexsort((fi[keep], -fs[keep]))[:k]
        assert np.array_equal(mi[q].cpu().numpy(), fi[keep][order])
        assert np.array_equal(ms[q].cpu().numpy(), fs[keep][order])


def test_dense_scorer_matches_torch_matmul(gauss):
    from mevi_b200.document_encoder import DocumentEncoder

    enc = DocumentEncoder()
    P = dev(gauss.X[:1024])
    q = dev(gauss.Q[0])
    out = enc.generate(q, p_reps=P).scores
    assert tuple(out.shape) == (1024,)
    ref = torch.matmul(torch.tensor(gauss.Q[0]), torch.tensor(gauss.X[:1024]).transpose(0, 1))  # document_encoder.py:132
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-4)
    out2 = enc.compute_similarity(dev(gauss.Q[:5]), P)
    np.testing.assert_allclose(out2.cpu().numpy(), gauss.Q[:5] @ gauss.X[:1024].T, rtol=1e-5, atol=1e-4)
    assert torch.equal(enc.compute_similarity(dev(gauss.Q[:3]), P[:3], bmm=True), torch.sum(dev(gauss.Q[:3]) * P[:3], -1))


@pytest.mark.parametrize("nb", [10, 100])
def test_rerank_all_candidates_sorted_like_the_shipped_recipe(case, nb):
    """`--save_hard_neg <corpus size>` (marco_eval_nci_rq.sh:26): the reference writes EVERY candidate, sorted
    (main_models.py:4012-4014, 4046-4053).  rerank_all returns them as a CSR; hn_lines_all formats them."""
    from mevi_b200.rerank import ClusterIndex, ClusterReranker, hn_lines_all

    dec = case.load(f"beam{nb}_labels.npy")
    clus = case.pickle("rqclus.pkl")
    rr = ClusterReranker(dev(case.X), ClusterIndex.from_codes(case.codes, case.K), mode="stream")
    off, ids, scores = rr.rerank_all(case.Q, dec)
    off, ids, scores = off.cpu().numpy(), ids.cpu().numpy(), scores.cpu().numpy()
    ref = oracle.cluster_rerank(case.Q, case.X, clus, dec, topk=None)
    X64, Q64 = case.X.astype(np.float64), case.Q.astype(np.float64)
    for q, (d_, s_, nd) in enumerate(ref):
        a, b = off[q], off[q + 1]
        assert b - a == nd == len(d_)
        np.testing.assert_allclose(scores[a:b], s_, rtol=1e-5, atol=1e-4)
        assert (np.diff(scores[a:b]) <= 0).all() and sorted(ids[a:b].tolist()) == sorted(d_.tolist())
        swapped = np.nonzero(ids[a:b] != d_)[0]  # order may differ only between (near-)equal scores
        for p_ in swapped:
            assert abs(float(X64[ids[a + p_]] @ Q64[q]) - float(X64[d_[p_]] @ Q64[q])) <= 1e-4 + 1e-5 * abs(s_[p_])
    # text format: byte-equal to the reference's writer on the same (ids, scores), with the [:save_hard_neg] slices
    texts = [f"query {i}" for i in range(len(ref))]
    for shn in (None, 7, 10 ** 7):
        lines = hn_lines_all(texts, torch.from_numpy(off), torch.from_numpy(ids), torch.from_numpy(scores), shn, None)
        for q in range(len(ref)):
            a, b = off[q], off[q + 1] if shn is None else min(off[q + 1], off[q] + shn)
            assert lines[q] == oracle.hn_result_line(texts[q], "", ids[a:b], scores[a:b])
    # the k best of the full list are what the top-k kernel returns
    s_k, i_k, _ = rr.rerank(case.Q, dec, topk=50)
    for q in range(len(ref)):
        kk = min(50, off[q + 1] - off[q])
        assert np.array_equal(s_k[q, :kk].cpu().numpy(), scores[off[q] : off[q] + kk])


@pytest.mark.parametrize("aggr", ["add", "max"])
def test_rerank_all_multiclus_aggregation(gauss, aggr):
    """`--doc_multiclus > 1` (main_models.py:3998-4011): a document listed under several of the query's leaves appears
    once, its scores added or maximised."""
    from mevi_b200.rerank import ClusterIndex, ClusterReranker

    dec = gauss.load("beam10_labels.npy")
    clus = {k: list(v) for k, v in gauss.pickle("rqclus.pkl").items()}
    for q in range(dec.shape[0]):  # a document of the query's first non-empty leaf is also listed under its second one
        hit = [tuple(int(v) for v in d_) for d_ in dec[q] if tuple(int(v) for v in d_) in clus]
        if len(hit) >= 2 and hit[0] != hit[1]:
            clus[hit[1]] = sorted(set(clus[hit[1]]) | {clus[hit[0]][0]})
    rr = ClusterReranker(dev(gauss.X), ClusterIndex.from_cluster_dict(clus, gauss.K), mode="stream")
    off, ids, scores = rr.rerank_all(gauss.Q, dec, multiclus_score_aggr=aggr)
    off, ids, scores = off.cpu().numpy(), ids.cpu().numpy(), scores.cpu().numpy()
    ref = oracle.cluster_rerank(gauss.Q, gauss.X, clus, dec, topk=None, multiclus_aggr=aggr)
    n_dup = 0
    for q, (d_, s_, nd) in enumerate(ref):
        a, b = off[q], off[q + 1]
        n_dup += nd - len(d_)
        assert b - a == len(d_) and len(set(ids[a:b].tolist())) == b - a
        np.testing.assert_allclose(scores[a:b], s_, rtol=1e-5, atol=2e-4)
        assert sorted(ids[a:b].tolist()) == sorted(d_.tolist())
    assert n_dup > 0  # the test data did contain documents under two of a query's leaves


def _hn_rows(text):
    rows = []
    for line in text.rstrip("\n").split("\n"):
        q, gt, docs, scores = line.split("\t")
        rows.append((q, gt, [int(v) for v in docs.split(",")] if docs else [],
                     np.array([float(v) for v in scores.split(",")] if scores else [], dtype=np.float64)))
    return rows


@pytest.mark.parametrize("case_name,nb", [("gauss768", 100), ("small64", 10)])
def test_product_equals_the_executed_reference_loop(case_name, nb):
    """The kernels against the output of main_models.py:3912-4055 EXECUTED verbatim (tests/golden/make_rerank_golden.py):
    the full sorted candidate list of the shipped recipe, the hard-negative lines' scores and ground-truth field, the
    --knn_topk_by_step lists (= the leading pool_size of the sort), and both --doc_multiclus aggregations."""
    import os
    import pickle

    from conftest import GOLDEN, golden_case
    from mevi_b200.rerank import ClusterIndex, ClusterReranker, hn_lines_all

    case = golden_case(case_name)
    dec = case.load(f"beam{nb}_labels.npy")
    load = lambda v: pickle.load(open(os.path.join(GOLDEN, "rerank", f"{case_name}_{v}.pkl"), "rb"))
    fx = load("shipped")
    rr = ClusterReranker(dev(case.X), ClusterIndex.from_codes(case.codes, case.K), mode="stream")
    off, ids, scores = (t.cpu().numpy() for t in rr.rerank_all(case.Q, dec))
    want = _hn_rows(fx["lines"])
    X64, Q64 = case.X.astype(np.float64), case.Q.astype(np.float64)
    rs = np.random.RandomState(11)
    gt = [[int(rs.randint(case.X.shape[0]))] for _ in range(len(case.Q))]
    gt_scores = ctx().dense_scores(dev(case.Q), dev(case.X[[g[0] for g in gt]])).cpu().numpy()
    for q in range(len(case.Q)):
        a, b = off[q], off[q + 1]
        assert b - a == fx["ndoc"][q] == len(fx["docs"][q])
        np.testing.assert_allclose(scores[a:b], want[q][3], rtol=1e-5, atol=1e-4)
        assert sorted(ids[a:b].tolist()) == sorted(fx["docs"][q])
        for p_ in np.nonzero(ids[a:b] != np.array(fx["docs"][q]))[0]:  # order may differ only between (near-)equal scores
            assert abs(float(X64[ids[a + p_]] @ Q64[q]) - float(X64[fx["docs"][q][p_]] @ Q64[q])) <= 1e-4
        np.testing.assert_allclose(gt_scores[q, q], float(want[q][1]), rtol=1e-5, atol=1e-4)
    # the line writer on the product's own lists reproduces the reference's text field by field (scores to 1e-5)
    texts = [f"query {i}" for i in range(len(case.Q))]
    gts = [str(float(np.float32(gt_scores[q, q]))) for q in range(len(case.Q))]
    lines = hn_lines_all(texts, torch.from_numpy(off), torch.from_numpy(ids), torch.from_numpy(scores), case.X.shape[0], gts)
    for q, (ln, w) in enumerate(zip(_hn_rows("\n".join(lines) + "\n"), want)):
        assert ln[0] == w[0] and len(ln[2]) == len(w[2])
    # --knn_topk_by_step 1, pool_size 50
    tk = load("topk_by_step")
    s_k, i_k, n_k = rr.rerank(case.Q, dec, topk=50)
    i_k, n_k = i_k.cpu().numpy(), n_k.cpu().numpy()
    for q in range(len(case.Q)):
        kk = len(tk["docs"][q])
        assert n_k[q] == tk["ndoc"][q] and kk == min(50, tk["ndoc"][q]) and (i_k[q, kk:] == -1).all()
        for p_ in np.nonzero(i_k[q, :kk] != np.array(tk["docs"][q]))[0]:
            assert abs(float(X64[i_k[q, p_]] @ Q64[q]) - float(X64[tk["docs"][q][p_]] @ Q64[q])) <= 1e-4
    # --doc_multiclus 2
    mc = load("multiclus_dict")
    rr2 = ClusterReranker(dev(case.X), ClusterIndex.from_cluster_dict(mc, case.K), mode="stream")
    for aggr in ("add", "max"):
        fm = load(f"multiclus_{aggr}")
        wm = _hn_rows(fm["lines"])
        off2, ids2, sc2 = (t.cpu().numpy() for t in rr2.rerank_all(case.Q, dec, multiclus_score_aggr=aggr))
        for q in range(len(case.Q)):
            a, b = off2[q], off2[q + 1]
            assert sorted(ids2[a:b].tolist()) == sorted(fm["docs"][q])
            np.testing.assert_allclose(sc2[a:b][:200], wm[q][3], rtol=1e-5, atol=2e-4)
