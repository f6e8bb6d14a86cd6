"""Host-side logic that needs no GPU: sharding rule, dictionaries, text formats, constructor surface."""
import numpy as np
import pytest
import torch

from oracle import oracle


def test_shard_bounds_is_the_reference_rule():
    from mevi_b200.dist_utils import shard_bounds

    for n in (0, 1, 7, 8841823, 21015324):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            per = n // world  # pq.py:219-224
            for r, (s, e) in enumerate(spans):
                assert s == per * r and e == (n if r + 1 == world else s + per)
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_codes_to_dicts_matches_reference_pickles(case):
    from mevi_b200.pq import codes_to_dicts

    clus, mapping = codes_to_dicts(case.codes, 0, True)
    ref_clus, ref_map = case.pickle("rqclus.pkl"), case.pickle("rqmapping.pkl")
    assert clus == ref_clus and mapping == ref_map
    assert list(clus) == list(ref_clus) and list(mapping) == list(ref_map)
    import pickle

    assert pickle.dumps(clus) == pickle.dumps(ref_clus) and pickle.dumps(mapping) == pickle.dumps(ref_map)
    # offset start (sharded ranks)
    c2, m2 = codes_to_dicts(case.codes[100:200], 100, True)
    assert all(m2[i] == ref_map[i] for i in range(100, 200))


def test_constructor_surface_and_out_of_scope_branches():
    from mevi_b200.pq import ProductQuantization

    pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
    assert tuple(pq.codebook.shape) == (4, 32, 768) and pq.codebook.dtype == torch.float32
    assert pq.subvector_cents == 32 and pq.last_dim == 768 and pq.get_preds is False
    assert "codebook" in pq.state_dict()
    # iptol2 cannot run in the reference either (self.extracol is never created, pq.py:113-117);
    # tied NCI centroids need the T5 lm_head
    for bad in (dict(pq_type="rq", dist_mode="iptol2"), dict(pq_type="rq", tie_nci_pq_centroid=1)):
        with pytest.raises(NotImplementedError):
            ProductQuantization(**bad)


def test_beam_search_mirror_matches_reference_golden(case):
    from mevi_b200.pq import ProductQuantization

    pq = ProductQuantization("rq", case.M, case.meta["bits"], "l2", case.d, "kmeans", "grad")
    with torch.no_grad():
        pq.codebook.copy_(torch.tensor(case.codebook))
    for nb in (10, 100):
        lab, sc = pq.beam_search(torch.tensor(case.Q), nb, return_proba=True)
        assert lab.dtype == torch.int64
        assert (lab.numpy() == case.load(f"beam{nb}_labels.npy")).all()
        np.testing.assert_allclose(sc.numpy(), case.load(f"beam{nb}_scores.npy"), rtol=1e-6)


def test_forward_mirror_matches_oracle(gauss):
    from mevi_b200.pq import ProductQuantization

    pq = ProductQuantization("rq", gauss.M, 5, "l2", gauss.d, "kmeans", "grad")
    with torch.no_grad():
        pq.codebook.copy_(torch.tensor(gauss.codebook))
    proba, index, loss = pq.forward(torch.tensor(gauss.X[:256]))
    assert tuple(proba.shape) == (256, gauss.M, gauss.K) and loss is None
    assert (index.numpy() == gauss.codes[:256]).all()
    rec = pq.get_reconstruct_vector(index)
    ref = sum(torch.tensor(gauss.codebook[j])[index[:, j]] for j in range(gauss.M))
    assert torch.allclose(rec, ref)


def test_kmeanspp_seeding_is_seeded_and_spread():
    from mevi_b200.trainer import kmeanspp_init

    rs = np.random.RandomState(0)
    centers = rs.standard_normal((8, 16)) * 10
    x = (centers[rs.randint(0, 8, 2000)] + rs.standard_normal((2000, 16)) * 0.1).astype(np.float32)
    a = kmeanspp_init(x, 8, np.random.RandomState(41)).numpy()
    b = kmeanspp_init(torch.from_numpy(x), 8, np.random.RandomState(41)).numpy()
    assert np.array_equal(a, b) and a.dtype == np.float32
    # one seed per well-separated blob
    owner = ((a[:, None, :] - centers[None]) ** 2).sum(-1).argmin(1)
    assert len(set(owner.tolist())) == 8


def test_text_formats_match_reference_writers(tmp_path):
    from mevi_b200 import faiss_search, rerank

    dists = np.array([[1.5, 0.1], [2.25, -3.0]], dtype=np.float32)
    ids = np.array([[7, 3], [1, -1]], dtype=np.int64)
    qf = tmp_path / "q.tsv"
    qf.write_text("what is x\t12\nwho is y\t13\n")
    out = tmp_path / "o.txt"
    faiss_search.to_file(str(qf), str(out), dists, ids)
    lines = out.read_text().splitlines()
    assert lines[0] == "what is x\t\t7,3\t1.5,0.10000000149011612"
    assert lines[0] == oracle.faiss_result_line("what is x", ids[0], dists[0])
    hn = rerank.hn_lines(["what is x", "who is y"], dists, ids)
    assert hn[0] == oracle.hn_result_line("what is x", "", ids[0], dists[0])
    assert hn[1] == "who is y\t\t1\t2.25"  # padding (-1) is not printed
    arr = np.arange(12, dtype=np.float32).reshape(3, 4)
    p = tmp_path / "emb.bin"
    arr.tofile(p)
    assert np.array_equal(faiss_search.read(str(p), 4), arr)


def test_grouped_rerank_plan_covers_every_pair_tile_exactly_once():
    """Host planning of the leaf-grouped re-rank (mevi_b200/rerank.py): tiles never straddle a leaf, every
    (query, leaf) pair meets every tile of its leaf exactly once, in the round its preceding candidate count puts it."""
    import torch

    from mevi_b200.rerank import GROUP_COLS, build_leaf_tiles, plan_grouped_rounds

    rs = np.random.RandomState(0)
    sizes = rs.randint(1, 700, size=50)
    sizes[3], sizes[7], sizes[9], sizes[11] = 1, 128, 129, 3000
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
    row0, nrows, lt0, src = build_leaf_tiles(off)
    assert torch.equal(torch.sort(src[src >= 0]).values.long(), torch.arange(int(off[-1])))
    for t in range(row0.numel()):
        a, b = int(row0[t]), int(row0[t]) + int(nrows[t])
        leaf = int(torch.searchsorted(off, torch.tensor(a), right=True)) - 1
        assert off[leaf] <= a and b <= off[leaf + 1] and 1 <= b - a <= 128
    nq, L = 300, 20
    ql = torch.from_numpy(np.stack([rs.choice(50, size=L, replace=False) for _ in range(nq)]).astype(np.int32))
    ql[5, 3] = -1
    ql[8, :] = -1
    seen = []
    for r, (it, ig, gq) in enumerate(plan_grouped_rounds(off, lt0, ql, (2000,))):
        gq = gq.view(-1, GROUP_COLS)
        for i in range(it.numel()):
            t, g = int(it[i]), int(ig[i])
            leaf = int(np.searchsorted(lt0.numpy(), t, side="right")) - 1
            seen += [(r, q, leaf, t) for q in gq[g][gq[g] >= 0].tolist()]
    exp = set()
    for q in range(nq):
        cum = 0
        for j in range(L):
            leaf = int(ql[q, j])
            if leaf < 0:
                continue
            exp |= {(0 if cum < 2000 else 1, q, leaf, t) for t in range(int(lt0[leaf]), int(lt0[leaf + 1]))}
            cum += int(sizes[leaf])
    assert len(seen) == len(set(seen)) and set(seen) == exp


def test_grouped_rerank_tile_plan_covers_every_pair_tile_exactly_once():
    """The default plan (plan_grouped_tile_rounds): round 0 = first tile of the leading `boot_leaves` leaves of a query,
    round 1 = everything else; every (query, leaf) pair meets every tile of its leaf exactly once, and in round 1 the
    further tiles of a leaf meet ALL its queries in full groups."""
    import torch

    from mevi_b200.rerank import GROUP_COLS, build_leaf_tiles, plan_grouped_tile_rounds

    rs = np.random.RandomState(1)
    sizes = rs.randint(1, 700, size=40)
    sizes[2], sizes[5], sizes[6] = 1, 128, 129
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
    _, _, lt0, _ = build_leaf_tiles(off)
    nq, L, BL = 200, 12, 5
    ql = torch.from_numpy(np.stack([rs.choice(40, size=L, replace=False) for _ in range(nq)]).astype(np.int32))
    ql[4, 2] = -1
    ql[9, :] = -1
    plan = plan_grouped_tile_rounds(lt0, ql, BL)
    assert len(plan) == 2
    plan3 = plan_grouped_tile_rounds(lt0, ql, (2, BL))  # the sample grown in two steps: same pairs, first tiles split
    assert len(plan3) == 3
    assert sum(int((p[2] >= 0).sum()) for p in plan3[:2]) == int((plan[0][2] >= 0).sum())
    assert torch.equal(plan3[2][0], plan[1][0]) and torch.equal(plan3[2][2], plan[1][2])
    seen = []
    for r, (it, ig, gq) in enumerate(plan):
        gq = gq.view(-1, GROUP_COLS)
        assert it.dtype == torch.int32 and ig.dtype == torch.int32 and (it.numel() == 0 or int(ig.max()) < gq.shape[0])
        for i in range(it.numel()):
            t, g = int(it[i]), int(ig[i])
            leaf = int(np.searchsorted(lt0.numpy(), t, side="right")) - 1
            seen += [(r, q, leaf, t) for q in gq[g][gq[g] >= 0].tolist()]
    exp = set()
    for q in range(nq):
        for j in range(L):
            leaf = int(ql[q, j])
            if leaf < 0:
                continue
            t0, t1 = int(lt0[leaf]), int(lt0[leaf + 1])
            exp.add((0 if j < BL else 1, q, leaf, t0))
            exp |= {(1, q, leaf, t) for t in range(t0 + 1, t1)}
    assert len(seen) == len(set(seen)) and set(seen) == exp
    # per-query rows of the bootstrap fit the candidate buffers by construction
    assert BL * 128 <= 8192
    # wide items (a tile meets up to 4 consecutive groups of its leaf: first group | count << 24) cover the same set
    wide = plan_grouped_tile_rounds(lt0, ql, BL, maxg_sample=2, maxg_last=4)
    seen_w = []
    for r, (it, ig, gq) in enumerate(wide):
        gq = gq.view(-1, GROUP_COLS)
        assert torch.equal(gq, plan[r][2].view(-1, GROUP_COLS))  # the groups themselves do not change
        for i in range(it.numel()):
            t, g0, ng = int(it[i]), int(ig[i]) & 0xFFFFFF, int(ig[i]) >> 24
            assert 1 <= ng <= (2 if r == 0 else 4) and g0 + ng <= gq.shape[0]
            leaf = int(np.searchsorted(lt0.numpy(), t, side="right")) - 1
            for g in range(g0, g0 + ng):
                seen_w += [(r, q, leaf, t) for q in gq[g][gq[g] >= 0].tolist()]
    assert len(seen_w) == len(set(seen_w)) and set(seen_w) == exp
    assert wide[1][0].numel() < plan[1][0].numel()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU arithmetic, oracle port) must print ONE JSON line with the
    keys the driver reads, without touching CUDA."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-sample", "2048"], capture_output=True, text=True, timeout=300,
                         env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "rq_encode_docs_per_sec" and j["unit"] == "docs/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_oracle_flat_search_partition_form_equals_sort_form():
    rs = np.random.RandomState(1)
    Q = rs.standard_normal((32, 48)).astype(np.float32)
    Dm = rs.standard_normal((9000, 48)).astype(np.float32)
    for docs, k, block in ((Dm, 100, 2048), (Dm[:50], 100, 65536), (Dm, 1000, 4096)):
        a = oracle.flat_ip_topk(Q, docs, k, block=block)
        b = oracle.flat_ip_topk_partition(Q, docs, k, block=block)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
