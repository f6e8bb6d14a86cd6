"""Full-size (BASELINE.json configs[1]: 8,841,823 x 768) checks through size-independent properties —
the oracle cannot finish this size in seconds, so parity here is carried by invariants of the path:

* shard invariance: encoding the corpus in the reference's row blocks (pq.py:218-225, ragged last block)
  gives the same codes as one call (what `get_document_cluster` over nrank ranks relies on);
* kernel-vs-kernel: the tensor path equals the direct fp32 kernel (the literal restatement of
  pq.py:124-131,300-305) on a strided sample, modulo flagged fp32 ties;
* reconstruction: X - sum_j codebook[j][code_j] equals the residual the kernel writes (pq.py:304-305), and the
  chosen centroid is the float64-nearest at every level on a 262k-row slice;
* Lloyd conservation: per-centroid counts sum to N, per-centroid sums add up to the column sums of X;
* inverted lists: every document lands in exactly one leaf, leaves ascending (SURVEY 8c invariants iii/iv).
"""
import numpy as np
import pytest
import torch

from oracle import oracle
from conftest import golden_case
from gpu_util import ctx

pytestmark = pytest.mark.gpu

N, D, M, K = 8841823, 768, 4, 32


@pytest.fixture(scope="module")
def corpus():
    gauss = golden_case("gauss768")
    free, _ = torch.cuda.mem_get_info(0)
    if free < 70 << 30:
        pytest.skip("needs ~60 GB of free HBM")
    g = torch.Generator(device="cuda:0")
    g.manual_seed(1234)
    X = torch.empty((N, D), dtype=torch.float32, device="cuda:0")
    for a in range(0, N, 1 << 20):
        X[a : a + (1 << 20)].normal_(generator=g)
    cb = torch.from_numpy(np.ascontiguousarray(gauss.codebook)).cuda(0)
    codes = ctx().rq_encode(X, cb, mode="auto")
    yield X, cb, codes
    del X, codes
    torch.cuda.empty_cache()


def _same_modulo_fp32_ties(Xs, cb, codes_a, codes_b, max_frac=1e-4):
    """Bit-equal codes, except rows the oracle's float64 arbiter classifies as fp32 ties (counted, bounded)."""
    bad = (codes_a != codes_b).any(dim=1).nonzero().flatten()
    if bad.numel():
        rep = oracle.classify_code_mismatches(Xs[bad].cpu().numpy(), cb.cpu().numpy(), codes_a[bad].cpu().numpy(),
                                              codes_b[bad].cpu().numpy())
        assert rep["n_hard"] == 0, rep
    assert bad.numel() <= max(1, max_frac * codes_a.shape[0])


def test_codes_in_range_and_level0_equals_kmeans_assignment(corpus):
    X, cb, codes = corpus
    assert codes.dtype == torch.int32 and tuple(codes.shape) == (N, M)
    assert int(codes.min()) >= 0 and int(codes.max()) < K
    # level 0 of the encode is the assignment step of Lloyd on codebook[0] (two entry points, one answer)
    buf = torch.empty(K * D + K, device="cuda:0")
    assign = torch.empty(N, dtype=torch.int32, device="cuda:0")
    ctx().kmeans_step(X, cb[0].contiguous(), buf, assign=assign)
    differ = (assign != codes[:, 0]).nonzero().flatten()
    assert differ.numel() <= 1e-5 * N
    if differ.numel():
        _same_modulo_fp32_ties(X[differ], cb[:1].contiguous(), assign[differ].view(-1, 1), codes[differ, :1].contiguous(), max_frac=1.0)


def test_shard_invariance_with_reference_row_blocks(corpus):
    from mevi_b200.dist_utils import shard_bounds

    X, cb, codes = corpus
    c = ctx()
    for nrank in (3, 8):
        for rank in range(nrank):
            a, b = shard_bounds(N, rank, nrank)
            part = c.rq_encode(X[a:b], cb, mode="auto")
            _same_modulo_fp32_ties(X[a:b], cb, part, codes[a:b])


def test_tensor_path_equals_direct_fp32_kernel_on_sample(corpus):
    X, cb, codes = corpus
    c = ctx()
    idx = torch.arange(0, N, 41, device="cuda:0")  # 215,654 rows spread over the whole corpus
    Xs = X[idx].contiguous()
    exact = c.rq_encode(Xs, cb, mode="exact")
    _same_modulo_fp32_ties(Xs, cb, exact, codes[idx])


def test_residual_is_x_minus_selected_centroids(corpus):
    X, cb, codes = corpus
    c = ctx()
    a, b = 4_000_000, 4_262_144
    res = torch.empty((b - a, D), device="cuda:0")
    part = c.rq_encode(X[a:b], cb, mode="exact", residual=res)
    want = X[a:b].clone()
    for j in range(M):  # same elementwise order as pq.py:304-305
        want -= cb[j][part[:, j].long()]
    assert torch.equal(res, want)
    # the oracle on a slice of the same rows: same codes, and the error shrinks level by level
    rows = X[a : a + 4096].cpu().numpy()
    ref = oracle.rq_encode(rows, cb.cpu().numpy())
    rep = oracle.classify_code_mismatches(rows, cb.cpu().numpy(), ref, codes[a : a + 4096].cpu().numpy())
    assert rep["n_hard"] == 0
    # argmin property in float64 on a larger slice: at every level the chosen centroid is (within fp32 resolution)
    # the nearest one to the running residual
    r = X[a:b].double()
    cb64 = cb.double()
    for j in range(M):
        d2 = torch.cdist(r, cb64[j]).pow(2)
        chosen = d2.gather(1, part[:, j].long().view(-1, 1)).squeeze(1)
        assert bool((chosen <= d2.min(1).values * (1 + 2e-6)).all())
        r = r - cb64[j][part[:, j].long()]


def test_lloyd_step_conserves_counts_and_sums(corpus):
    X, cb, _ = corpus
    c = ctx()
    buf = torch.empty(K * D + K, device="cuda:0")
    assign = torch.empty(N, dtype=torch.int32, device="cuda:0")
    c.kmeans_step(X, cb[0].contiguous(), buf, assign=assign)
    counts = buf[K * D :].double()
    assert int(counts.sum().item()) == N
    assert torch.equal(torch.bincount(assign.long(), minlength=K).double(), counts)
    col = torch.zeros(D, dtype=torch.float64, device="cuda:0")
    for a in range(0, N, 1 << 20):
        col += X[a : a + (1 << 20)].double().sum(0)
    got = buf[: K * D].view(K, D).double().sum(0)
    # fp32 running sums over up to ~7M rows per centroid: relative to the mass that was added
    scale = float(X[: 1 << 20].abs().double().sum(0).mean().item()) * (N / (1 << 20))
    assert float((got - col).abs().max().item()) <= 2e-6 * scale


def test_inverted_lists_partition_the_corpus(corpus):
    from mevi_b200.rerank import ClusterIndex

    _, _, codes = corpus
    index = ClusterIndex.from_codes(codes, K)
    off, ids = index.leaf_offsets, index.leaf_docids
    assert int(off[0]) == 0 and int(off[-1]) == N and bool((off[1:] > off[:-1]).all())  # only non-empty leaves
    assert torch.equal(torch.sort(ids.long()).values, torch.arange(N, device=ids.device))
    # ascending doc ids inside every leaf (the reference appends docs in id order, pq.py:236-242)
    inc = ids[1:] > ids[:-1]
    starts = torch.zeros(N - 1, dtype=torch.bool, device=ids.device)
    starts[(off[1:-1] - 1).long()] = True
    assert bool((inc | starts).all())
    # leaf key of a document == its code tuple
    key = torch.zeros(N, dtype=torch.int64, device=codes.device)
    for j in range(M):
        key = key * K + codes[:, j].long()
    leaf_of_pos = torch.repeat_interleave(torch.arange(index.n_leaves, device=ids.device), (off[1:] - off[:-1]))
    assert torch.equal(index.leaf_keys[leaf_of_pos], key[ids.long()])


def test_rerank_auto_mode_takes_the_grouped_path_and_equals_the_streaming_kernel(corpus):
    """Default ClusterReranker at the bench shape: 2,048 queries x 100 leaves share leaves (~9 pairs per leaf), so the
    leaf-grouped tensor path runs; its answer must be the streaming kernel's (ids modulo fp32 score ties)."""
    from mevi_b200.pq import ProductQuantization
    from mevi_b200.rerank import ClusterIndex, ClusterReranker

    X, cb, codes = corpus
    free, _ = torch.cuda.mem_get_info(0)
    if free < 60 << 30:
        pytest.skip("needs ~50 GB more HBM for the leaf-ordered copy and its fp16 image")
    g = torch.Generator(device="cuda:0")
    g.manual_seed(4321)
    nq, L, k = 2048, 100, 100
    Q = torch.empty((nq, D), device="cuda:0").normal_(generator=g)
    pq = ProductQuantization("rq", M, 5, "l2", D, "kmeans", "grad")
    with torch.no_grad():
        pq.codebook.copy_(cb.cpu())
    dec = torch.cat([pq.beam_search(Q[a : a + 128], L) for a in range(0, nq, 128)])
    rr = ClusterReranker(X, ClusterIndex.from_codes(codes, K))
    assert rr.mode == "auto" and rr._grouped is not None
    s_g, i_g, n_g = rr.rerank(Q, dec, topk=k)
    assert rr.last_path == "grouped"
    rr.mode = "stream"
    saved, rr._grouped = rr._grouped, None
    s_s, i_s, n_s = rr.rerank(Q, dec, topk=k)
    rr._grouped = saved
    assert rr.last_path == "stream" and torch.equal(n_g, n_s)
    assert torch.equal(s_g, s_s)  # exact fp32 re-score in the same summation order
    differ = (i_g != i_s)
    assert float(differ.float().mean()) < 1e-3
    # where ids differ the two documents tie in score
    if bool(differ.any()):
        q_idx, pos = differ.nonzero(as_tuple=True)
        a = (X[i_g[q_idx, pos]] * Q[q_idx]).sum(1)
        b = (X[i_s[q_idx, pos]] * Q[q_idx]).sum(1)
        assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max())
    # a few queries that share few leaves stay on the streaming kernel
    rr.mode = "auto"
    rr.rerank(Q[:64], dec[:64], topk=k)
    assert rr.last_path == "stream"
    del rr
    torch.cuda.empty_cache()


def test_streaming_rerank_equals_float64_numpy_on_sampled_queries(corpus):
    """Independent check of the streaming re-rank kernel at the full corpus size (the grouped-vs-streaming test above
    compares two of this repo's own paths): for a few queries the candidate rows of the decoded leaves are pulled to the
    host and scored in float64 numpy, sorted by (score desc, id asc) as main_models.py:3915-4014 does.  Ids must agree
    except where the float64 scores tie within fp32 resolution; scores within 1e-5 relative (north star: 1e-3)."""
    import numpy as np

    from mevi_b200.pq import ProductQuantization
    from mevi_b200.rerank import ClusterIndex, ClusterReranker

    X, cb, codes = corpus
    g = torch.Generator(device="cuda:0")
    g.manual_seed(97)
    nq, L, k = 6, 100, 100
    Q = torch.empty((nq, D), device="cuda:0").normal_(generator=g)
    pq = ProductQuantization("rq", M, 5, "l2", D, "kmeans", "grad")
    with torch.no_grad():
        pq.codebook.copy_(cb.cpu())
    dec = pq.beam_search(Q, L)
    index = ClusterIndex.from_codes(codes, K)
    rr = ClusterReranker(X, index, mode="stream")
    s, i, ncand = rr.rerank(Q, dec, topk=k)
    assert rr.last_path == "stream"
    ql = index.lookup(dec).cpu().numpy()
    off = index.leaf_offsets.cpu().numpy()
    for q in range(nq):
        leaves = [int(l) for l in ql[q] if l >= 0]
        ids = torch.cat([index.leaf_docids[off[l] : off[l + 1]] for l in leaves]).long()
        assert int(ncand[q]) == ids.numel() and ids.numel() > 1000
        rows = X[ids].cpu().numpy().astype(np.float64)
        sc = rows @ Q[q].cpu().numpy().astype(np.float64)
        ids_np = ids.cpu().numpy()
        order = np.lexsort((ids_np, -sc))[:k]
        got_i, got_s = i[q].cpu().numpy(), s[q].cpu().numpy().astype(np.float64)
        assert np.allclose(got_s, sc[order], rtol=1e-5, atol=1e-5)
        bad = np.nonzero(got_i != ids_np[order])[0]
        for p in bad:  # a different document at this rank: the two must tie at fp32 resolution
            mine = sc[np.nonzero(ids_np == got_i[p])[0][0]]
            assert abs(mine - sc[order][p]) <= 4e-6 * max(1.0, abs(mine)), (q, p)
        assert bad.size <= 2
    del rr
    torch.cuda.empty_cache()
