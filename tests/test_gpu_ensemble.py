"""Device path of the result fusion (csrc/ensemble.cu through the C ABI) against
  * the report text of the UNMODIFIED reference scripts (tests/golden/ensemble/*, made by make_ensemble_golden.py) and
  * the python-dictionary restatement of ensemble_marco.py:181-191, 221-240 on adversarial lists (duplicates whose
    later score overwrites the earlier one, exact score ties, -1 padding, ragged lengths, repeated leaves)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["marco", "nqdpr"])
def test_device_report_text_equals_reference(kind, tmp_path):
    import make_ensemble_golden as g

    import mevi_b200
    from mevi_b200 import ensemble

    work = str(tmp_path / "in")
    g.ensemble_inputs(work)
    ofile = str(tmp_path / "report.txt")
    args = g.marco_args(work, ofile) if kind == "marco" else g.nq_args(work, ofile)
    args.device = "cuda"
    before = mevi_b200.get_context(0).launches
    (ensemble.combine_main_marco if kind == "marco" else ensemble.combine_main_nqdpr)(args)
    assert mevi_b200.get_context(0).launches > before, "no kernel of the library ran"
    golden = open(os.path.join(HERE, "golden", "ensemble", f"ensemble_{kind}_report.txt")).read()
    assert open(ofile).read() == golden
    # the rank caches hold what the reference's python loop computes
    import pickle

    ranks, num = pickle.load(open(os.path.join(work, "coarse_cr4gt.pkl"), "rb"))
    mapping = pickle.load(open(os.path.join(work, "rqmapping.pkl"), "rb"))
    template = {"query": 0, "pred": 2, "score": 3}
    if kind == "nqdpr":
        template["_by_line"] = True
    preds, _, _ = ensemble.parse_file(os.path.join(work, "ance.txt"), template)
    _, _, coarse = ensemble.parse_file(os.path.join(work, "coarse.tsv"), {"query": 0, "cluster": 1},
                                       os.path.join(work, "ance.txt") if kind == "nqdpr" else None)
    want, wnum = ensemble.cluster_rankings(preds, coarse, mapping)
    assert num == wnum and ranks == want


def _random_lists(rs, nq, P, ndocs, L, M, K):
    codes = rs.randint(0, K, size=(ndocs, M)).astype(np.int32)
    mapping = {d: tuple(int(v) for v in codes[d]) for d in range(ndocs)}
    preds, scores, coarse = {}, {}, {}
    pool_scores = np.round(rs.rand(64) * 4, 1)  # few distinct values: many exact ties after fusion
    distinct = None
    for q in range(nq):
        n = int(rs.randint(0, P + 1)) if q % 5 else P
        ids = rs.randint(0, ndocs, size=n)
        if n > 4:
            ids[rs.randint(0, n, size=n // 4)] = ids[rs.randint(0, n, size=n // 4)]  # duplicates
            ids[rs.randint(0, n, size=max(n // 10, 1))] = -1  # padding ids (also duplicated)
        preds[q] = [int(v) for v in ids]
        scores[q] = [float(v) for v in rs.choice(pool_scores, size=n)]
        # ordered leaf list with exactly L - 2 distinct leaves: two repeats whose LAST index counts
        uniq = sorted(set(mapping.values()))
        lv = [uniq[i] for i in rs.choice(len(uniq), size=L - 2, replace=False)]
        lv.insert(int(rs.randint(1, L - 2)), lv[0])
        lv.append(lv[2])
        coarse[q] = [list(t) for t in lv]
        distinct = len(set(lv))
    return codes, mapping, preds, scores, coarse, distinct


@pytest.mark.parametrize("nq,P,L", [(97, 37, 10), (64, 2000, 100), (16, 4096, 20), (5, 1, 4)])
def test_fusion_kernels_vs_python_dictionaries(nq, P, L):
    from mevi_b200.ensemble import DeviceFusion, cluster_rankings, fuse, ranking_of

    rs = np.random.RandomState(nq * 1000 + P)
    M, K = 3, 8
    codes, mapping, preds, scores, coarse, distinct = _random_lists(rs, nq, P, 500, L, M, K)
    fusion = DeviceFusion(list(preds), preds, scores)
    ranks, num = fusion.cluster_ranks(coarse, codes)
    want_ranks, want_num = cluster_rankings(preds, coarse, mapping)
    assert num == want_num == distinct
    assert ranks == want_ranks
    for alpha, beta, gamma in [(0.6, 0.03, 0.02), (0.2, 0.1, 0.5), (0.0, 1.0, 0.0), (1.5, 0.0, 0.7)]:
        ranked, fused, counts = fusion.fuse(alpha, beta, gamma, num)
        ranked, fused, counts = ranked.cpu().numpy(), fused.cpu().numpy(), counts.cpu().numpy()
        for i, q in enumerate(preds):
            want = fuse(preds[q], scores[q], want_ranks[q], alpha, beta, gamma, num)
            order = ranking_of(want)
            assert counts[i] == len(order)
            assert ranked[i, :counts[i]].tolist() == order, (q, alpha, beta, gamma)
            # float64, bit for bit
            assert fused[i, :counts[i]].tobytes() == np.array([want[p] for p in order], dtype=np.float64).tobytes()
            assert (ranked[i, counts[i]:] == -1).all() and np.isneginf(fused[i, counts[i]:]).all()
    # evaluator look-ups
    truth = {q: [int(v) for v in rs.randint(-1, 500, size=rs.randint(0, 5))] for q in preds}
    ranked, _, counts = fusion.fuse(0.6, 0.03, 0.02, num)
    pos = fusion.positions(ranked, counts, truth)
    raw = fusion.positions(fusion.ids, fusion.count, truth)
    for q in preds:
        order = ranking_of(fuse(preds[q], scores[q], want_ranks[q], 0.6, 0.03, 0.02, num))
        assert pos[q] == [order.index(g) if g in order else None for g in truth[q]]
        assert raw[q] == [preds[q].index(g) if g in preds[q] else None for g in truth[q]]  # first occurrence
    inv = [sorted(set(rs.choice(nq, size=rs.randint(0, 3)).tolist())) for _ in range(500)]
    offsets = np.zeros(501, dtype=np.int32)
    offsets[1:] = np.cumsum([len(v) for v in inv])
    array = np.array([x for v in inv for x in v] + [0], dtype=np.int32)
    hits = fusion.first_hits(ranked, counts, offsets, array)
    for q in preds:
        order = ranking_of(fuse(preds[q], scores[q], want_ranks[q], 0.6, 0.03, 0.02, num))
        want = next((j for j, res in enumerate(order) if q in array[offsets[res]:offsets[res + 1]]), None)
        assert hits[q] == want


def test_missing_document_is_a_key_error():
    from mevi_b200.ensemble import DeviceFusion, mapping_to_codes

    mapping = {0: (0, 0), 2: (1, 1)}
    codes = mapping_to_codes(mapping)
    assert codes.shape == (3, 2)
    for bad in (1, 7):
        fusion = DeviceFusion(["q"], {"q": [0, bad, -1]}, {"q": [1.0, 2.0, 3.0]})
        with pytest.raises(KeyError):
            fusion.cluster_ranks({"q": [[0, 0], [1, 1]]}, codes)


def test_device_epilogue_of_the_rerank_output():
    """Fusion straight from device results: float32 re-rank scores widened to float64 on the device, codes from the
    encoder — the lists never visit the host."""
    import mevi_b200
    from mevi_b200.ensemble import fuse, ranking_of

    c = mevi_b200.get_context(0)
    g = torch.Generator(device="cpu").manual_seed(5)
    nq, k, n, M, K, L = 33, 100, 5000, 4, 4, 12
    codes = torch.randint(0, K, (n, M), generator=g, dtype=torch.int32)
    ids = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(nq)])
    sc = torch.rand((nq, k), generator=g).sort(dim=1, descending=True).values
    leaves = torch.stack([codes[torch.randperm(n, generator=g)[:L]] for _ in range(nq)])
    cranks, num = c.ensemble_cluster_ranks(ids.cuda(), None, codes.cuda(), leaves.cuda())
    num = num.cpu().numpy()
    ranked, fused, counts = c.ensemble_fuse(ids.cuda(), sc.cuda().double(), cranks, None, 0.6, 0.03, 0.02, int(num[0]))
    cr = cranks.cpu().numpy()
    for q in range(nq):
        lv = {}
        for i, t in enumerate(leaves[q].tolist()):
            lv[tuple(t)] = i
        if len(lv) != num[0]:
            continue  # a query whose random leaves collide has another distinct count; covered by the test above
        want_cr = [lv.get(tuple(codes[p].tolist()), len(lv)) for p in ids[q].tolist()]
        assert cr[q].tolist() == want_cr
        want = fuse(ids[q].tolist(), [float(v) for v in sc[q].double()], want_cr, 0.6, 0.03, 0.02, len(lv))
        assert ranked[q, :counts[q]].cpu().tolist() == ranking_of(want)
