"""Doc-sharded grouped re-rank at the bench shape on N GPUs: time per call with and without threshold sharing.
torchrun --nproc-per-node N tools/dist_rerank_share.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.dist_utils import shard_bounds
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ctx = mevi_b200.get_context(dev.index)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().to(dev)
n = 8841823
s, e = shard_bounds(n, rank, world)
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
X = torch.empty((e - s, 768), device=dev)
for a in range(0, e - s, 1 << 20): X[a:a + (1 << 20)].normal_(generator=g)
codes = ctx.rq_encode(X, cb)
g.manual_seed(4321)
Q = torch.empty((6980, 768), device=dev).normal_(generator=g)
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb.cpu())
dec = torch.cat([pq.beam_search(Q[a:a + 1024], 100) for a in range(0, 6980, 1024)])
LEAF = os.environ.get("LEAF_PARTITION", "0") != "0"
if LEAF:
    index, X = ClusterIndex.from_sharded_codes(X, codes, 32, s, device_index=dev.index)
else:
    index = ClusterIndex.from_codes(codes, 32, id_base=s, device_index=dev.index)
D_leaf = ctx.gather_rows(X, index.leaf_docids)
del X
rr = ClusterReranker(None, index, mode="grouped", D_leaf=D_leaf)
if rank == 0: print("leaf-partitioned" if LEAF else "row blocks", "rows", D_leaf.shape[0], "leaves", index.n_leaves, flush=True)
ref = None
for share, boot in ((False, (8, 63)), (True, (8, 63)), (False, (16, 100)), (True, (16, 100))):
    rr.SHARE_THRESHOLDS = share
    rr.BOOT_LEAVES = boot
    for _ in range(2): out = rr.rerank(Q, dec, topk=100)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(8): out = rr.rerank(Q, dec, topk=100)
    torch.cuda.synchronize(); dist.barrier(); ms = (time.perf_counter() - t0) / 8 * 1e3
    if ref is None: ref = out
    same = float((out[1] == ref[1]).float().mean())
    if rank == 0: print(f"world {world} share {share} boot {boot}: {ms:.2f} ms per call, path {rr.last_path}, weak {rr.last_weak_queries} failed {rr.last_failed_queries}, ids equal to the first run {same:.6f}", flush=True)
dist.destroy_process_group()
