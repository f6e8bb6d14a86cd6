"""Multi-GPU functional check (NCCL, one process per GPU): sharded Lloyd training, doc-sharded flat search with
all-gather merge, doc-sharded cluster re-rank.  Launch: torchrun --nproc-per-node 2 tests/dist_check.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import mevi_b200
from mevi_b200 import faiss_search
from mevi_b200.dist_utils import shard_bounds
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker
from oracle import oracle

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ctx = mevi_b200.get_context(dev.index)
rs = np.random.RandomState(0)
n, d = 40001, 768
centers = rs.standard_normal((64, d)).astype(np.float32)
X = (centers[rs.randint(0, 64, n)] + 0.5 * rs.standard_normal((n, d))).astype(np.float32)
Q = rs.standard_normal((50, d)).astype(np.float32)

# 1. sharded training through the drop-in entry point: all ranks end with the same codebook; rank 0 gets all codes
pq = ProductQuantization("rq", 4, 5, "l2", d, "kmeans", "grad"); pq.device_index = dev.index; pq.lloyd_iters = 8
pq.initialize(None, X, rank, 41, None, 1024)
cb = pq.codebook.detach().clone().to(dev)
gathered = [torch.empty_like(cb) for _ in range(world)]
dist.all_gather(gathered, cb)
assert all(torch.equal(g, gathered[0]) for g in gathered), "codebooks differ across ranks"
if rank == 0:
    codes = pq.last_preds
    assert codes.shape == (n, 4)
    rep = oracle.classify_code_mismatches(X, pq.codebook.detach().numpy(), oracle.rq_encode(X, pq.codebook.detach()), codes)
    mse = oracle.quantisation_mse(X, pq.codebook.detach().numpy(), codes)
    print(f"[train] world={world} mse={mse:.5f} hard mismatches={rep['n_hard']} ties={rep['n_ties']}", flush=True)
    assert rep["n_hard"] == 0
codes_all = torch.from_numpy(pq.last_preds).to(dev) if rank == 0 else torch.empty((n, 4), dtype=torch.int32, device=dev)
dist.broadcast(codes_all, 0)

# 2. doc-sharded flat search == single-GPU search
dists, idx = faiss_search.search(Q, X, d, 100, "Flat", device_index=dev.index)
if rank == 0:
    s1, i1 = ctx.flat_ip_topk(torch.from_numpy(Q).to(dev), torch.from_numpy(X).to(dev), 100)
    same = (i1.cpu().numpy() == idx).mean()
    print(f"[flat] sharded vs single ids equal {same:.4f}, max score diff {np.abs(s1.cpu().numpy() - dists).max():.2e}", flush=True)
    assert same > 0.995 and np.allclose(s1.cpu().numpy(), dists, rtol=1e-5, atol=2e-4)

# 3. doc-sharded re-rank == single-GPU re-rank
dec = pq.beam_search(torch.from_numpy(Q).to(dev), 20)
s, e = shard_bounds(n, rank, world)
index = ClusterIndex.from_codes(codes_all[s:e], 32, id_base=s, device_index=dev.index)
rr = ClusterReranker(torch.from_numpy(X[s:e]).to(dev), index)
sc, ids, nc = rr.rerank(Q, dec, topk=100)
if rank == 0:
    full = ClusterIndex.from_codes(codes_all, 32, device_index=dev.index)
    ql = full.lookup(dec)
    s0, i0, n0 = ctx.cluster_rerank(torch.from_numpy(Q).to(dev), torch.from_numpy(X).to(dev), full.leaf_offsets, full.leaf_docids, ql, 100)
    same = (i0 == ids).float().mean().item()
    print(f"[rerank] sharded vs single ids equal {same:.4f}, candidates equal {bool((n0 == nc).all())}", flush=True)
    assert same > 0.995 and bool((n0 == nc).all()) and torch.allclose(s0, sc, rtol=1e-5, atol=2e-4)
# 4. the same through the leaf-grouped tensor path (thresholds shared between the ranks after every round)
rrg = ClusterReranker(torch.from_numpy(X[s:e]).to(dev), index, mode="grouped")
rrg.BOOTSTRAP_MIN, rrg.BOOT_LEAVES, rrg.SHARE_THRESHOLDS = 128, (2, 5), True
sc_g, ids_g, nc_g = rrg.rerank(Q, dec, topk=100)
assert rrg.last_path.startswith("grouped") and rrg._share_now, (rrg.last_path, rrg._share_now)
if rank == 0:
    same = (i0 == ids_g).float().mean().item()
    print(f"[rerank grouped] sharded vs single ids equal {same:.4f}, path {rrg.last_path}", flush=True)
    assert same > 0.995 and bool((n0 == nc_g).all()) and torch.allclose(s0, sc_g, rtol=1e-5, atol=2e-4)
# 5. leaf-partitioned index: every leaf moved whole to one rank (one all-to-all of the rows), same answers
index_l, X_own = ClusterIndex.from_sharded_codes(torch.from_numpy(X[s:e]).to(dev), codes_all[s:e], 32, s, device_index=dev.index)
tot = torch.tensor([X_own.shape[0]], device=dev); dist.all_reduce(tot)
assert int(tot.item()) == n and index_l.doc_ids.numel() == X_own.shape[0]
sizes_l = torch.tensor([index_l.n_leaves], device=dev); dist.all_reduce(sizes_l)
for mode in ("stream", "grouped"):
    rrl = ClusterReranker(X_own, index_l, mode=mode)
    rrl.BOOTSTRAP_MIN, rrl.BOOT_LEAVES = 128, (2, 5)
    sc_l, ids_l, nc_l = rrl.rerank(Q, dec, topk=100)
    if rank == 0:
        same = (i0 == ids_l).float().mean().item()
        print(f"[rerank leaf-partitioned {mode}] vs single ids equal {same:.4f}, path {rrl.last_path}, leaves over all ranks "
              f"{int(sizes_l.item())} (single index {full.n_leaves})", flush=True)
        assert int(sizes_l.item()) == full.n_leaves, "a leaf lives on more than one rank"
        assert same > 0.995 and bool((n0 == nc_l).all()) and torch.allclose(s0, sc_l, rtol=1e-5, atol=2e-4)
dist.barrier()
if rank == 0: print("DIST CHECK OK", flush=True)
dist.destroy_process_group()
