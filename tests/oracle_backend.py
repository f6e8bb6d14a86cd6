"""A CPU stand-in for `mevi_b200._lib.Context`, built on the oracle — TEST DOUBLE ONLY.
It lets the world_size-2 gloo tests drive the product's trainer / merge sequencing (sharding,
collectives, update order) on a machine without a GPU.  It never ships: the product package does
not import it."""
import numpy as np
import torch

from oracle import oracle


class OracleBackend:
    torch_device = torch.device("cpu")

    def kmeans_step(self, R, centroids, sums_counts, assign=None, assign_stride=1, inertia=None, mode="auto"):
        K, d = centroids.shape
        a, sums, counts, inert, _ = oracle.lloyd_step(R.numpy(), centroids.numpy())
        sums_counts[: K * d] = torch.from_numpy(sums.astype(np.float32).ravel())
        sums_counts[K * d :] = torch.from_numpy(counts.astype(np.float32))
        if assign is not None:
            assign.copy_(torch.from_numpy(a))
        if inertia is not None:
            inertia[0] = inert

    def kmeans_update(self, sums_counts, centroids, n_empty=None):
        K, d = centroids.shape
        sums = sums_counts[: K * d].view(K, d)
        counts = sums_counts[K * d :]
        nz = counts > 0
        centroids[nz] = sums[nz] / counts[nz].unsqueeze(1)
        if n_empty is not None:
            n_empty[0] = int((~nz).sum())

    def residual_update(self, R, centroids, assign, assign_stride=1):
        R -= centroids[assign.long()]

    def topk_merge(self, scores_in, ids_in):
        S, nq, k = scores_in.shape
        s = scores_in.permute(1, 0, 2).reshape(nq, S * k).numpy()
        i = ids_in.permute(1, 0, 2).reshape(nq, S * k).numpy()
        out_s = np.full((nq, k), -np.inf, np.float32)
        out_i = np.full((nq, k), -1, np.int64)
        for q in range(nq):
            keep = i[q] >= 0
            order = np.lexsort((i[q][keep], -s[q][keep]))[:k]
            out_s[q, : len(order)] = s[q][keep][order]
            out_i[q, : len(order)] = i[q][keep][order]
        return torch.from_numpy(out_s), torch.from_numpy(out_i)
