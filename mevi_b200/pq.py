"""Drop-in `ProductQuantization` for the RQ path of MEVI/pq.py, backed by libmevi_b200.

Same constructor, method names, argument meaning, side effects (`self.codebook`,
`self.last_preds`, `self.get_preds`) and on-disk formats as the reference
(MEVI/pq.py:15-741): `pq_type in ('rq','pq','opq')` (the shipped scripts use
'rq'), `dist_mode in ('l2','ip')`, `pq_init_method in ('none','kmeans','avg')`,
`pq_update_method in ('grad','kmeans','ema','fixpq',...)`.  The hot arithmetic
runs in hand-written CUDA kernels through the C ABI (include/mevi_b200.h);
PyTorch only owns memory, streams and the process group.  Branches that need
packages or state the reference itself does not have here raise
NotImplementedError instead of silently doing something else: faiss index
import/export (`pq_init_method='faiss'`, `pq_update_method='faiss'`, `codebook_from_index`,
`build_faiss_index`, `unsupervised_update_codebook_faiss` — and with them the only way the
reference obtains an OPQ rotation), tied NCI centroids (T5 lm_head), `do_sample` beam search, and
`dist_mode='iptol2'`, which the reference cannot run either — it writes to
`self.extracol`, an attribute that is never created (pq.py:113-117), so every
iptol2 path dies with AttributeError.

Differences that are deliberate:
  * codebook training is full-batch Lloyd, data-parallel over the row blocks of
    pq.py:218-225, with ONE all-reduce of the fused [K*d+K] sums|counts buffer
    per iteration, instead of sklearn MiniBatchKMeans on rank 0
    (pq.py:449,557-563).  Codebooks therefore differ from sklearn's; encode
    parity is always defined given a codebook.
  * every rank takes part in `initialize` when torch.distributed is initialised
    (the reference parks ranks 1.. behind the broadcast at pq.py:483-484).
"""
from __future__ import annotations

import os
import os.path as osp
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from . import _lib
from .dist_utils import shard_bounds


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized()


class ProductQuantization(nn.Module):
    # MEVI/pq.py:16-80
    def __init__(
        self,
        pq_type: str = "pq",
        subvector_num: int = 32,
        subvector_bits: int = 8,
        dist_mode: str = "ip",
        emb_size: int = 768,
        pq_init_method: str = "faiss",
        pq_update_method: str = "grad",
        tie_nci_pq_centroid: int = 0,
        lm_head: nn.Parameter = None,
        centroid_update_loss: str = "none",
        rq_topk_score: str = "prod",
    ):
        super().__init__()
        assert pq_type in ("pq", "opq", "rq")
        assert dist_mode in ("ip", "l2", "iptol2")
        if dist_mode == "iptol2":
            raise NotImplementedError(
                "dist_mode='iptol2': the reference writes to self.extracol, which it never creates (pq.py:113-117), "
                "so this mode raises AttributeError there as well")
        if tie_nci_pq_centroid:
            raise NotImplementedError("tie_nci_pq_centroid couples the codebook to the T5 lm_head: out of scope")
        self.pq_type = pq_type
        self.subvector_num = subvector_num
        self.subvector_bits = subvector_bits
        self.subvector_cents = 2 ** subvector_bits
        self.dist_mode = dist_mode
        self.emb_size = emb_size
        self.pq_init_method = pq_init_method
        self.pq_update_method = pq_update_method
        self.tie_nci_pq_centroid = tie_nci_pq_centroid
        self.centroid_update_loss = centroid_update_loss
        self.rq_topk_score = rq_topk_score
        self.get_preds = False
        self.last_dim = emb_size if pq_type == "rq" else emb_size // subvector_num  # pq.py:50-54
        # pq.py:67-68 — a CPU float Parameter [M, K, last_dim]
        self.codebook = nn.Parameter(
            torch.empty(subvector_num, self.subvector_cents, self.last_dim), requires_grad=(pq_update_method == "grad")
        ).type(torch.FloatTensor)
        if pq_type == "opq":  # pq.py:69-71
            self.rotate = nn.Parameter(torch.empty(emb_size, emb_size), requires_grad=False).type(torch.FloatTensor)
        if pq_update_method == "ema":  # pq.py:72-80
            self.decay = 0.99
            self.eps = 1e-5
            self.restart_unused_codes = True
            self.register_buffer("cluster_size_ema", torch.zeros(subvector_num, self.subvector_cents))
            self.register_buffer("embed_ema", self.codebook.detach().clone())
        # knobs of the B200 trainer / encoder (not in the reference signature)
        self.kernel_mode = "auto"  # 'auto' | 'exact' | 'tensor'
        self.lloyd_iters = 25
        self.lloyd_tol = 1e-7
        self.init_sample = 16384
        self.device_index: Optional[int] = None

    # ------------------------------------------------------------------ #
    def _ctx(self) -> "_lib.Context":
        return _lib.get_context(self.device_index)

    def rq_minus_centroids(self, embeddings, centroids):  # pq.py:121-122
        return embeddings - centroids[..., : self.last_dim]

    def compute_scores(self, a, b):  # pq.py:124-131 (API compatibility; the kernels do this on the device)
        if self.dist_mode == "ip":
            result = a * b
        else:
            result = -((a - b) ** 2)
        return torch.sum(result, dim=-1)

    def get_codebook(self):  # pq.py:133-141
        return self.codebook

    def fix(self):  # pq.py:435-438
        self.codebook.requires_grad_(False)
        if self.pq_type == "opq":
            self.rotate.requires_grad_(False)

    # ------------------------------------------------------------------ #
    # encode                                                             #
    # ------------------------------------------------------------------ #
    @torch.no_grad()
    def get_rq_document_cluster(self, doc_embeddings, cluster: torch.Tensor, start: int, ending: int, rank: int,
                                batch_size: int = 1024):
        """pq.py:281-305.  Fills `cluster` (int32 [ending-start, M]) in place.
        `doc_embeddings` may be an np.ndarray / np.memmap (streamed host->device
        in pinned chunks, pq.py:283's shard copy never materialises on the host)
        or a CUDA tensor (encoded in place on the device).  `batch_size` is
        accepted for signature compatibility; batching is the kernel's business."""
        ctx = self._ctx()
        cb = self.get_codebook().detach()
        if isinstance(doc_embeddings, torch.Tensor) and doc_embeddings.is_cuda:
            x = doc_embeddings[start:ending].contiguous().float()
            codes = ctx.rq_encode(x, cb.to(x.device).contiguous(), metric=self.dist_mode, mode=self.kernel_mode)
            cluster.copy_(codes.to(cluster.device))
            return
        part = np.asarray(doc_embeddings[start:ending]) if not isinstance(doc_embeddings, np.memmap) else doc_embeddings[start:ending]
        if part.dtype != np.float32 or not part.flags.c_contiguous:
            part = np.ascontiguousarray(part, dtype=np.float32)
        assert cluster.dtype == torch.int32 and cluster.is_contiguous() and not cluster.is_cuda
        cb_host = np.ascontiguousarray(cb.cpu().numpy(), dtype=np.float32)
        self.last_encode_stats = ctx.rq_encode_host(part, cb_host, cluster.numpy(), metric=self.dist_mode,
                                                    mode=self.kernel_mode)

    @torch.no_grad()
    def get_pq_document_cluster(self, doc_embeddings, cluster: torch.Tensor, start: int, ending: int, rank: int,
                                batch_size: int = 1024):
        """pq.py:249-279: per sub-vector nearest centroid ('opq': after x @ rotate.T, 259-261).  Host rows are
        streamed to the device in 256k-row pieces; the encode is `mevi_pq_encode`, the rotation a plain fp32
        library GEMM.  Fills `cluster` (int32 [ending-start, M]) in place."""
        ctx = self._ctx()
        dev = torch.device("cuda", ctx.device)
        cb = self.get_codebook().detach().to(dev).contiguous()
        rot_t = self.rotate.detach().to(dev).T.contiguous() if self.pq_type == "opq" else None
        # `mevi_pq_encode` itself takes the tensor route when the shape allows it (M*K <= 128: the RQ kernel on a
        # block-padded codebook, sub-vector kernel as the arbiter of flagged rows; csrc/pq_encode.cu)
        step = 1 << 18
        for a in range(start, ending, step):
            b = min(a + step, ending)
            if isinstance(doc_embeddings, torch.Tensor):
                x = doc_embeddings[a:b].to(device=dev, dtype=torch.float32).contiguous()
            else:
                x = torch.from_numpy(np.ascontiguousarray(doc_embeddings[a:b], dtype=np.float32)).to(dev)
            if rot_t is not None:
                x = _matmul_fp32(x, rot_t)
            codes = ctx.pq_encode(x, cb, metric=self.dist_mode)
            cluster[a - start : b - start].copy_(codes.to(cluster.device))

    @torch.no_grad()
    def get_document_cluster(self, doc_embeddings, rank: int, nrank: int, batch_size: int = 1024,
                             return_mapping: bool = False):
        """pq.py:216-247: row block of this rank, encode, then the
        {code tuple -> [doc ids]} / {doc id -> code tuple} dictionaries."""
        num_docs = doc_embeddings.shape[0]
        start, ending = shard_bounds(num_docs, rank, nrank)
        cluster = torch.empty((ending - start, self.subvector_num), dtype=torch.int32)
        func = self.get_rq_document_cluster if self.pq_type == "rq" else self.get_pq_document_cluster  # pq.py:228-231
        func(doc_embeddings, cluster, start, ending, rank, batch_size)
        doc_cluster, new_mapping = codes_to_dicts(cluster.numpy(), start, return_mapping)
        print("Number of document clusters:", len(doc_cluster))
        if return_mapping:
            return doc_cluster, new_mapping
        return doc_cluster

    @torch.no_grad()
    def get_document_cluster_simple(self, return_mapping: bool = False):
        """pq.py:200-214: dictionaries from the codes the trainer kept in `last_preds`."""
        assert self.get_preds
        cluster, mapping = codes_to_dicts(np.asarray(self.last_preds), 0, True)
        del self.last_preds
        self.get_preds = False
        if return_mapping:
            return cluster, mapping
        return cluster

    # ------------------------------------------------------------------ #
    # build                                                              #
    # ------------------------------------------------------------------ #
    @torch.no_grad()
    def initialize(self, index_file, doc_emb, rank, seed, pq_cluster_path, encode_batch_size):
        """pq.py:440-486."""
        if self.pq_init_method == "none":
            return
        if self.pq_init_method == "faiss":
            raise NotImplementedError("pq_init_method='faiss' (faiss ResidualQuantizer import) is out of scope")
        use_file = index_file is not None and osp.isfile(index_file)
        if self.pq_init_method == "avg":  # pq.py:471-482
            if rank == 0:
                if use_file:
                    self.codebook.copy_(torch.load(index_file, map_location="cpu"))
                else:
                    torch.nn.init.normal_(self.codebook.data, mean=0.0, std=0.01)
                print(f"Intializing codebook using average document embedding after {use_file} use file...")
                self.init_pq_using_document_cluster(doc_emb, pq_cluster_path, encode_batch_size)
            if _dist_on():
                self._broadcast_codebook()
            return
        assert self.pq_init_method.endswith("kmeans")
        if not use_file:
            self.get_preds = True
        if use_file:
            if rank == 0:
                print("Intializing codebook with torch file...")
                tensor = torch.load(index_file, map_location="cpu")
                self.codebook.copy_(tensor)
                del tensor
            if _dist_on():
                self._broadcast_codebook()
        else:
            print("Intializing codebook by kmeans clustering...")
            self.unsupervised_update_codebook_manually(doc_emb, seed, self.pq_init_method)
            if rank == 0 and index_file is not None:
                torch.save(self.codebook, index_file)  # pq.py:469-470: the Parameter itself

    def _broadcast_codebook(self):
        # pq.py:483-484; NCCL needs a device buffer
        if dist.get_backend() == "nccl":
            buf = self.codebook.data.cuda()
            dist.broadcast(buf, 0)
            self.codebook.data.copy_(buf.cpu())
        else:
            dist.broadcast(self.codebook.data, 0)

    @torch.no_grad()
    def unsupervised_update_codebook(self, doc_emb, rank, seed, align=False):
        """pq.py:526-542."""
        if self.pq_update_method == "faiss":
            raise NotImplementedError("pq_update_method='faiss' is out of scope")
        if self.pq_update_method.endswith("kmeans"):
            self.get_preds = True
            ori_codebook = self.codebook.detach().clone() if align else None
            self.unsupervised_update_codebook_manually(doc_emb, seed, self.pq_update_method)
            if align:  # pq.py:540-541
                self.align_codebook(ori_codebook)

    @torch.no_grad()
    def unsupervised_update_codebook_manually(self, doc_emb, seed, kmeans_method):
        """pq.py:550-598, rq branch: per level k-means on the current residual,
        then residual -= centers[pred] (skipped after the last level, 591-593);
        sets `self.codebook` and `self.last_preds` (codes of every row, on rank 0).

        Full-batch Lloyd on the device, sharded over ranks (see module docstring)."""
        print("Updating codebook using KMeans...")
        assert self.pq_type != "opq"  # pq.py:553
        if kmeans_method != "kmeans":
            raise NotImplementedError(f"kmeans_method={kmeans_method!r} (pq.py:564-565 raises too)")
        from .trainer import train_pq_lloyd, train_rq_lloyd

        if self.pq_type == "pq":  # pq.py:568-581
            codebook, codes_all = train_pq_lloyd(
                doc_emb, M=self.subvector_num, K=self.subvector_cents, seed=int(seed), iters=self.lloyd_iters,
                tol=self.lloyd_tol, init_sample=self.init_sample, mode=self.kernel_mode, device_index=self.device_index)
            self.last_preds = codes_all
            self.last_train_info = getattr(train_pq_lloyd, "last_info", None)
            with torch.no_grad():
                self.codebook.copy_(codebook.cpu())
            return
        codebook, codes_all = train_rq_lloyd(
            doc_emb,
            M=self.subvector_num,
            K=self.subvector_cents,
            seed=int(seed),
            iters=self.lloyd_iters,
            tol=self.lloyd_tol,
            init_sample=self.init_sample,
            mode=self.kernel_mode,
            device_index=self.device_index,
            metric=self.dist_mode,
        )
        self.last_preds = codes_all  # np.int32 [N, M] on rank 0 (None elsewhere)
        self.last_train_info = getattr(train_rq_lloyd, "last_info", None)
        with torch.no_grad():
            self.codebook.copy_(codebook.cpu())

    # ------------------------------------------------------------------ #
    # leaf producer                                                      #
    # ------------------------------------------------------------------ #
    @torch.no_grad()
    def beam_search(self, doc_emb: torch.Tensor, num_return_sequences, num_beams=None, do_sample=False,
                    return_proba=False):
        """pq.py:613-713.  CUDA input, rq: `mevi_rq_beam_search` (one pass for the x.c table, then table
        arithmetic per level — csrc/beam.cu).  CPU input (the reference runs on `codebook.device`, the CPU by
        default) and the pq/opq branch: the reference's own tensor operations, restated below.
        Returns labels int64 [bs, beams, M] (+ beam scores)."""
        if num_beams is None:
            num_beams = num_return_sequences
        if do_sample:
            raise NotImplementedError("do_sample=True (torch.multinomial branch, pq.py:686-688) is out of scope")
        if self.pq_type == "rq" and doc_emb.is_cuda and self._beam_kernel_fits(int(num_beams), doc_emb.shape[-1]):
            ctx = _lib.get_context(doc_emb.device)
            cb = self.get_codebook().detach().to(doc_emb.device).contiguous()
            labels, scores = ctx.rq_beam_search(doc_emb.contiguous().float(), cb, int(num_beams), metric=self.dist_mode,
                                                prod=(self.rq_topk_score == "prod"))
            labels = labels.long()  # pq.py: the int32 seed column is promoted by cat with int64 codes
            return (labels, scores) if return_proba else labels
        # shapes beyond the kernel's shared-memory state (e.g. 8-bit codebooks with 100 beams: 25,600 candidates per
        # level) and the pq/opq branch run the reference's tensor-op formulation on the input's device
        return self._beam_search_tensor_ops(doc_emb, num_beams, return_proba)

    def _beam_kernel_fits(self, num_beams: int, d: int) -> bool:
        """The limits `mevi_rq_beam_search` enforces (csrc/beam.cu): M*K <= 2048, num_beams*K <= 16384 (rounded up to a
        power of two) and a per-query state below 200 KB of shared memory."""
        M, K = self.subvector_num, self.subvector_cents
        cap = 1
        while cap < num_beams * K:
            cap <<= 1
        sel = 64
        while sel < 2 * num_beams and sel < 1024:
            sel <<= 1
        smem = 8 * (M * K + 2 * num_beams) + 4 * (d + 2 * num_beams + cap) + 4 * (cap + 2 * num_beams * M) + 8 * sel
        return M * K <= 2048 and cap <= 16384 and smem <= 200 * 1024 and float(K) ** M >= num_beams

    def _beam_search_tensor_ops(self, doc_emb, num_beams, return_proba):
        """The reference's tensor-op formulation of pq.py:626-713, op for op (bit-identical to the reference on
        the CPU golden vectors, tests/test_host_logic.py)."""
        codebook = self.get_codebook().detach().to(doc_emb.device)
        rq = self.pq_type == "rq"
        if self.pq_type == "opq":  # pq.py:630-631
            doc_emb = torch.matmul(doc_emb, self.rotate.detach().to(doc_emb.device).T)
        K = self.subvector_cents
        bs = doc_emb.size(0)
        beam_scores = doc_emb.new_ones(bs, 1)
        temp_embed = doc_emb.unsqueeze(1).clone() if rq else None
        temp_index = torch.zeros((bs, 1, 1), device=doc_emb.device, dtype=torch.int32)
        for i in range(self.subvector_num):
            if rq:
                cur_codebook = codebook[i : i + 1].expand(bs, -1, -1).unsqueeze(1)
                proba = self.compute_scores(temp_embed.unsqueeze(-2), cur_codebook)
            else:  # pq.py:654-660
                cur_codebook = codebook[i].unsqueeze(0).expand(bs, -1, -1)
                cur_embed = doc_emb[:, i * self.last_dim : (i + 1) * self.last_dim].unsqueeze(1)
                proba = self.compute_scores(cur_embed, cur_codebook)
            proba = F.softmax(proba, dim=-1)
            if not rq:
                proba = beam_scores.unsqueeze(-1) * proba.unsqueeze(1)  # pq.py:669
            elif self.rq_topk_score == "prod":
                proba = beam_scores.unsqueeze(-1) * proba
            proba = proba.view(bs, -1)
            prev = beam_scores.size(1)
            beam_of = torch.div(torch.arange(prev * K, device=doc_emb.device), K, rounding_mode="floor")
            code_of = torch.arange(K, device=doc_emb.device).repeat(prev)
            if num_beams < proba.size(1):
                _, top = proba.topk(num_beams, dim=-1)
                prev_beams = beam_of[top].unsqueeze(-1)
                cur_code = code_of[top]
                beam_scores = proba.gather(1, top)
                temp_index = torch.cat(
                    [temp_index.gather(1, prev_beams.expand(-1, -1, temp_index.size(-1))), cur_code.unsqueeze(-1)], dim=-1)
                if rq and i != self.subvector_num - 1:
                    temp_embed = temp_embed.gather(1, prev_beams.expand(-1, -1, temp_embed.size(-1))) \
                        - codebook[i][cur_code][..., : self.last_dim]
            else:
                beam_scores = proba
                temp_index = torch.cat(
                    [temp_index.repeat_interleave(K, dim=1), code_of.unsqueeze(-1).unsqueeze(0).expand(bs, -1, -1)], dim=-1)
                if rq and i != self.subvector_num - 1:
                    temp_embed = temp_embed.repeat_interleave(K, dim=1) - codebook[i][code_of][..., : self.last_dim]
        assert beam_scores.size(1) == num_beams
        topk_label = temp_index[:, :, 1:]
        if return_proba:
            return topk_label, beam_scores
        return topk_label

    @torch.no_grad()
    def get_topk_document_mapping(self, doc_embeddings, rank: int, nrank: int, num_return_sequences: int,
                                  batch_size: int = 1024):
        """pq.py:715-741."""
        num_docs = doc_embeddings.shape[0]
        start, ending = shard_bounds(num_docs, rank, nrank)
        out = torch.empty((ending - start, num_return_sequences, self.subvector_num), dtype=torch.int32, device="cpu")
        dev = torch.device("cuda", self.device_index if self.device_index is not None else torch.cuda.current_device())
        for i in range(start, ending, batch_size):
            e = min(i + batch_size, ending)
            cur = torch.tensor(np.asarray(doc_embeddings[i:e]), device=dev)
            out[i - start : e - start] = self.beam_search(cur, num_return_sequences).to("cpu", torch.int32)
        return out

    # ------------------------------------------------------------------ #
    # training-time forward (tensor ops; not on the index hot path)       #
    # ------------------------------------------------------------------ #
    def forward(self, vecs, return_loss=True):
        """pq.py:307-319: (proba, index, loss); EMA codebook update when training with pq_update_method='ema'."""
        if self.pq_type == "rq":
            proba, index, loss = self.forward_rq(vecs, return_loss)
        else:
            proba, index, loss = self.forward_pq(vecs, return_loss)
        if self.training and self.pq_update_method == "ema":
            self.ema_update(vecs, index)
        return proba, index, loss

    def forward_pq(self, vecs, return_loss=True):
        """pq.py:321-337: (proba [B,M,K], index [B,M], loss)."""
        if self.pq_type == "opq":
            vecs = torch.matmul(vecs, self.rotate.T)
        vecs = vecs.view(vecs.size(0), self.subvector_num, -1)
        codebook = self.get_codebook().unsqueeze(0).expand(vecs.size(0), -1, -1, -1)
        proba = self.compute_scores(vecs.unsqueeze(-2), codebook)
        index = proba.max(dim=-1)[1]
        loss = None
        if return_loss and self.centroid_update_loss == "reconstruct":  # pq.py:329-333
            reconstruct_emb = self.get_reconstruct_vector(index).view(index.shape[0], self.subvector_num, self.last_dim)
            loss = ((vecs - reconstruct_emb) ** 2).mean()
        return proba, index, loss

    def forward_rq(self, vecs, return_loss=True):
        """pq.py:339-369: (proba [B,M,K], index [B,M], loss).  Like the reference, `vecs -= centroid`
        (pq.py:357) subtracts IN PLACE: the caller's tensor holds the residual before the last level afterwards —
        `ema_update` (called from `forward` with the same tensor) relies on it."""
        allproba, index = [], []
        codebook = self.get_codebook()
        use_rec = self.centroid_update_loss == "reconstruct"
        errors = []
        for i in range(self.subvector_num):
            cur_codebook = codebook[i : i + 1].expand(vecs.size(0), -1, -1)
            proba = self.compute_scores(vecs.unsqueeze(-2), cur_codebook)
            part_index = proba.max(dim=-1)[1]
            allproba.append(proba)
            index.append(part_index)
            cur_centroid = codebook[i][part_index]
            if use_rec:
                errors.append(vecs.detach() - cur_centroid)
            if i != self.subvector_num - 1:
                vecs -= cur_centroid.detach()
        proba = torch.stack(allproba, dim=1)
        index = torch.stack(index, dim=1)
        loss = (torch.stack(errors) ** 2).mean() if (return_loss and use_rec) else None
        return proba, index, loss

    # ------------------------------------------------------------------ #
    # EMA codebook update                                                 #
    # ------------------------------------------------------------------ #
    @torch.no_grad()
    def ema_sums_counts(self, vectors: torch.Tensor, idxs: torch.Tensor):
        """pq.py:373-393: per (level, centroid) sums of the vectors assigned to it and the assignment counts,
        as `(vectors_sum_per_cluster [M,K,last_dim], cluster_size [M,K])`.  The reference builds a one-hot
        matrix and a bmm; here each level is one `mevi_accumulate_by_code` pass (deterministic shared-memory
        accumulators, csrc/kmeans.cu).  For 'rq' every level sums the SAME `vectors` (pq.py:375-377 expands the
        input across levels), for 'pq'/'opq' level j sums sub-vector j."""
        M, K, w = self.subvector_num, self.subvector_cents, self.last_dim
        ctx = _lib.get_context(vectors.device)
        if self.pq_type == "opq":
            vectors = torch.matmul(vectors, self.rotate.to(vectors.device).T)
        vectors = vectors.detach().float().contiguous()
        codes = idxs.reshape(-1, M).to(torch.int32).contiguous()
        sums = torch.empty((M, K, w), dtype=torch.float32, device=vectors.device)
        counts = torch.empty((M, K), dtype=torch.float32, device=vectors.device)
        buf = torch.empty(K * w + K, dtype=torch.float32, device=vectors.device)
        for j in range(M):
            x = vectors if self.pq_type == "rq" else vectors[:, j * w : (j + 1) * w].contiguous()
            ctx.accumulate_by_code(x, codes[:, j], K, buf, assign_stride=M)
            sums[j].copy_(buf[: K * w].view(K, w))
            counts[j].copy_(buf[K * w :])
        return sums, counts

    @torch.no_grad()
    def ema_update(self, vectors, idxs):
        """pq.py:371-433.  Sums/counts on the device (`ema_sums_counts`), ONE all-reduce of each when
        torch.distributed is up (pq.py:395-397), then the reference's EMA / restart / normalisation steps on the
        [M,K,*] buffers."""
        M, K, w = self.subvector_num, self.subvector_cents, self.last_dim
        vectors_sum_per_cluster, cluster_size = self.ema_sums_counts(vectors, idxs)
        if _dist_on():
            dist.all_reduce(vectors_sum_per_cluster, op=dist.ReduceOp.SUM)
            dist.all_reduce(cluster_size, op=dist.ReduceOp.SUM)
        dev = self.cluster_size_ema.device
        self.cluster_size_ema.mul_(self.decay).add_(cluster_size.to(dev), alpha=1 - self.decay)
        self.embed_ema.mul_(self.decay).add_(vectors_sum_per_cluster.to(dev), alpha=1 - self.decay)
        if self.restart_unused_codes:  # pq.py:404-423
            if self.pq_type == "rq":
                temp = vectors.detach().unsqueeze(1).expand(-1, M, -1)
            else:
                v = torch.matmul(vectors, self.rotate.to(vectors.device).T) if self.pq_type == "opq" else vectors
                temp = v.detach().reshape(-1, M, w)
            B = temp.shape[0]
            if B < K:
                n_repeats = (K + B - 1) // B
                std = temp.new_ones(w) * 0.01 / np.sqrt(w)
                temp = temp.repeat(n_repeats, 1, 1)
                temp = temp + torch.rand_like(temp) * std
            _vectors_random = torch.stack(
                [temp[torch.randperm(temp.shape[0], device=temp.device), i][:K] for i in range(M)])
            if _dist_on():
                dist.broadcast(_vectors_random, 0)
            _vectors_random = _vectors_random.to(dev)
            usage = (self.cluster_size_ema.unsqueeze(-1) >= 1).float()
            self.embed_ema.mul_(usage).add_(_vectors_random * (1 - usage))
            usage = usage.squeeze(-1)
            self.cluster_size_ema.mul_(usage)
            self.cluster_size_ema.add_(torch.ones_like(self.cluster_size_ema) * (1 - usage))
        n = self.cluster_size_ema.sum(dim=1, keepdim=True)  # pq.py:426-432
        normalized_cluster_size = n * (self.cluster_size_ema + self.eps) / (n + K * self.eps)
        self.codebook.data[:] = (self.embed_ema / normalized_cluster_size.unsqueeze(-1)).to(self.codebook.device)

    def get_reconstruct_vector(self, index, codebook=None):
        """pq.py:768-784: rq = sum of the selected centroids; pq = their concatenation ('opq': rotated back)."""
        if codebook is None:
            codebook = self.get_codebook()[..., : self.last_dim]
        assert index.dim() in (1, 2)
        M = self.subvector_num
        parts = [codebook[j][index[..., j]] for j in range(M)]
        vectors = torch.stack(parts, dim=-2)
        if self.pq_type == "rq":
            return vectors.sum(dim=-2)
        vectors = vectors.reshape(*index.shape[:-1], -1)
        if self.pq_type == "opq":
            vectors = torch.matmul(vectors, self.rotate)
        return vectors


    def get_reconstruct_loss_for_embeddings(self, embeddings, labels):
        """pq.py:743-766: mean squared reconstruction error of `embeddings` [B,d] under codes `labels` [B,M]
        ('rq': the running residual after every level, stacked; 'pq': against the concatenated centroids)."""
        assert labels.dim() == 2
        codebook = self.get_codebook()[..., : self.last_dim]
        vectors = torch.stack([codebook[j][labels[:, j]] for j in range(self.subvector_num)], dim=1)  # [B,M,last_dim]
        if self.pq_type == "pq":
            diff = embeddings - vectors.reshape(labels.shape[0], -1)
        elif self.pq_type == "rq":
            diffs, cur = [], embeddings
            for i in range(self.subvector_num):
                cur = cur - vectors[:, i, :]
                diffs.append(cur)
            diff = torch.stack(diffs, 1)
        else:
            assert False  # pq.py:761-762: the reference has no opq branch here either
        return (diff ** 2).mean()

    def get_reconstruct_vector_matrix_multiply(self, index):
        """pq.py:786-799: soft reconstruction, `index` [bs, M, K] weights over the centroids of every level."""
        bs = index.shape[0]
        codebook = self.get_codebook()[..., : self.last_dim]
        w = index.reshape(-1, self.subvector_cents).unsqueeze(1)
        cbx = codebook.unsqueeze(0).expand(bs, -1, -1, -1).reshape(-1, self.subvector_cents, self.last_dim)
        output = torch.bmm(w, cbx).squeeze(1).view(bs, self.subvector_num, self.last_dim)
        if self.pq_type == "rq":
            return torch.sum(output, dim=1)
        return output.view(bs, -1)

    @torch.no_grad()
    def align_codebook(self, ori_codebook):
        """pq.py:600-611: per level, permute the new centroids so that they line up with `ori_codebook` under the
        score-maximising assignment (scipy's Hungarian solver, as the reference)."""
        from scipy.optimize import linear_sum_assignment

        new_codebook = self.codebook.new(*self.codebook.shape)
        for ori, cur, new in zip(ori_codebook, self.codebook, new_codebook):
            scores = self.compute_scores(ori.unsqueeze(0), cur.unsqueeze(1)).cpu().numpy()
            assign = linear_sum_assignment(scores, maximize=True)
            for cid, oid in zip(*assign):
                new[oid] = cur[cid]
        self.codebook.copy_(new_codebook)

    @torch.no_grad()
    def init_pq_using_document_cluster(self, doc_emb, cluster, batch_size):
        """pq.py:488-524: codebook[i][k] = mean of the (residual) embeddings of the documents whose code at level i
        is k, taken from an existing `rqclus*.pkl` / `pqclus*.pkl`; 'rq' subtracts the mean before the next level.
        `cluster` is the pickle path (as in the reference) or the dictionary itself.  The per-code means are one
        `mevi_accumulate_by_code` pass per level on the device (fp32 sums / counts; the reference accumulates
        sum/ndocs in float64), the residual update is `mevi_residual_update`.  Codes absent from the dictionary keep
        their previous centroid, as in the reference."""
        import pickle

        assert self.dist_mode in ("l2",)
        assert self.pq_type in ("pq", "rq")
        if isinstance(cluster, (str, os.PathLike)):
            with open(cluster, "rb") as fr:
                cluster = pickle.load(fr)
        ctx = self._ctx()
        dev = torch.device("cuda", ctx.device)
        M, K, w = self.subvector_num, self.subvector_cents, self.last_dim
        n = doc_emb.shape[0]
        codes = np.full((n, M), -1, dtype=np.int32)
        for key, docs in cluster.items():
            codes[np.asarray(docs, dtype=np.int64)] = np.asarray(key, dtype=np.int32)
        member = torch.from_numpy((codes[:, 0] >= 0)).to(dev)
        rows = torch.nonzero(member).squeeze(1)
        X = torch.from_numpy(np.ascontiguousarray(np.asarray(doc_emb), dtype=np.float32)).to(dev)
        X = X[rows].contiguous() if rows.numel() != n else X
        cdev = torch.from_numpy(codes).to(dev)[rows].contiguous()
        buf = torch.empty(K * w + K, dtype=torch.float32, device=dev)
        for i in range(M):
            part = X if self.pq_type == "rq" else X[:, i * w : (i + 1) * w].contiguous()
            ctx.accumulate_by_code(part, cdev[:, i], K, buf, assign_stride=M)
            sums, counts = buf[: K * w].view(K, w), buf[K * w :]
            seen = counts > 0
            cent = self.codebook.data[i].to(dev).clone()
            cent[seen] = sums[seen] / counts[seen].unsqueeze(1)
            self.codebook.data[i].copy_(cent.cpu())
            if self.pq_type == "rq" and i != M - 1:
                ctx.residual_update(X, cent.contiguous(), cdev[:, i], assign_stride=M)

    # ---- faiss-backed branches of the reference: explicit scope cuts (faiss is not part of this path) ----------
    def augment_xb(self, xb, phi=None):  # pq.py:82-87 (iptol2 helper; numpy only)
        norms = np.sum((xb ** 2), axis=-1)
        if phi is None:
            phi = np.max(norms)
        return np.hstack((xb, np.sqrt(phi - norms)[..., np.newaxis]))

    def augment_xq(self, xq):  # pq.py:89-95
        if isinstance(xq, torch.Tensor):
            return torch.cat((xq, xq.new_zeros(*xq.shape[:-1], 1)), dim=-1)
        return np.concatenate((xq, np.zeros((*xq.shape[:-1], 1), dtype=xq.dtype)), axis=-1)

    def wrapped_augment_xb(self, xb, index=None):  # pq.py:97-119
        if self.dist_mode != "iptol2":
            return xb
        raise NotImplementedError("dist_mode='iptol2' (pq.py:113-117 writes to a self.extracol that is never created)")

    def codebook_from_index(self, index, index_file=None):  # pq.py:143-173
        raise NotImplementedError("codebook_from_index reads a faiss index (faiss.read_index / index.rq.codebooks): "
                                  "faiss import/export is out of scope; load a .pt codebook through initialize()")

    def build_faiss_index(self, doc_embeddings, save_path=None):  # pq.py:175-198
        raise NotImplementedError("build_faiss_index trains a faiss RQ/PQ/OPQ index: out of scope; use "
                                  "initialize(..., pq_init_method='kmeans') for the device trainer")

    def unsupervised_update_codebook_faiss(self, doc_emb, seed):  # pq.py:544-548
        raise NotImplementedError("pq_update_method='faiss' is out of scope")


def _matmul_fp32(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Plain fp32 library GEMM (cuBLAS, TF32 off) for the OPQ rotation of pq.py:259-261."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return torch.matmul(x, w).contiguous()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def codes_to_dicts(codes: np.ndarray, start: int = 0, return_mapping: bool = True):
    """The dictionaries of pq.py:236-242 from an int code table: python-int tuples, doc ids ascending inside a leaf,
    leaves in order of first appearance (the insertion order of the reference's defaultdict) - so `pickle.dumps` of
    the result is byte-identical to the reference's rqclus / rqmapping files.  No per-row interpreter loop: rows are
    grouped with one stable sort by leaf key, leaves are ordered by their smallest doc id, and python objects are made
    by `ndarray.tolist()`; the only python-level loop runs over the leaves."""
    codes = np.ascontiguousarray(codes)
    n, M = codes.shape
    if n == 0:
        return {}, ({} if return_mapping else None)
    base = int(codes.max()) + 1
    key = np.zeros(n, dtype=np.int64)
    for j in range(M):
        key = key * base + codes[:, j].astype(np.int64)
    order = np.argsort(key, kind="stable")                     # doc ids ascending inside a leaf
    skey = key[order]
    starts = np.flatnonzero(np.concatenate(([True], skey[1:] != skey[:-1])))
    ends = np.concatenate((starts[1:], [n]))
    first_doc = order[starts]                                  # a stable sort puts the leaf's smallest row first
    leaf_order = np.argsort(first_doc, kind="stable")          # first appearance
    if return_mapping:
        tuples = list(map(tuple, codes.tolist()))              # one distinct tuple object per row, like tuple(v.tolist())
        mapping = dict(zip(range(start, start + n), tuples))
        key_of = lambda g: tuples[first_doc[g]]                # the leaf's key IS its first row's tuple (pq.py:238-240)
    else:
        mapping = None
        firsts = codes[first_doc].tolist()
        key_of = lambda g: tuple(firsts[g])
    docs_sorted = order + start
    doc_cluster = {}
    for g_ in leaf_order.tolist():
        doc_cluster[key_of(g_)] = docs_sorted[starts[g_] : ends[g_]].tolist()
    return doc_cluster, mapping


def _cli():
    # MEVI/pq.py:802-827 keeps these flags; the reference's own __main__ reads an undefined
    # `args.emb_size` (pq.py:825) and dies with AttributeError before doing anything.  Here the same
    # flags build the RQ codebook on the device and save it (torch.save of the Parameter).
    import argparse

    parser = argparse.ArgumentParser()
    parser.add_argument("--embedding_path", type=str, required=True)
    parser.add_argument("--save_path", type=str, required=True)
    parser.add_argument("--dist_mode", type=str, default="l2")
    parser.add_argument("--pq_type", type=str, default="rq")
    parser.add_argument("--subvector_num", type=int, default=4)
    parser.add_argument("--subvector_bits", type=int, default=4)
    parser.add_argument("--dim", type=int, default=768)
    parser.add_argument("--seed", type=int, default=41)
    args = parser.parse_args()
    assert args.dist_mode in ("l2", "ip", "iptol2")
    assert args.pq_type in ("pq", "opq", "rq")
    doc_embeddings = np.memmap(args.embedding_path, dtype=np.float32, mode="r").reshape(-1, args.dim)
    pq = ProductQuantization(args.pq_type, args.subvector_num, args.subvector_bits, args.dist_mode, args.dim, "kmeans")
    pq.initialize(args.save_path, doc_embeddings, 0, args.seed, None, 1024)


if __name__ == "__main__":
    _cli()
