"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the hot path shards.

  * embedding extraction: images are independent -> split the batch, no data-path collective; `gather_rows` collects the
    (B, D) embeddings when one rank needs them all.
  * 1-NN: gallery row-sharded in contiguous blocks; every rank scores all queries against its shard, the per-rank
    (squared distance, global index) pairs are all-gathered (8-12 bytes per query per rank) and merged with
    lowest-index tie-break by hfr_knn_merge.

These helpers only move metadata / small result vectors and work with both the nccl (GPU) and gloo (CPU tests)
back ends; the compute stays in libhfr.so.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_rows(n: int, world_size: int, rank: int):
    """Contiguous row block [start, stop) of rank `rank` (the first n % world ranks get one extra row)."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def split_batch(x, group=None):
    """This rank's slice of a batch that every rank holds (data-parallel extraction)."""
    rank, ws = world(group)
    a, b = shard_rows(len(x), ws, rank)
    return x[a:b]


def shard_layout(n_local: int, y_local, group=None):
    """-> (global row offset of this rank's shard, labels of the whole gallery in global row order, rows per rank)."""
    rank, ws = world(group)
    y_local = np.asarray(y_local)
    if ws == 1:
        return 0, y_local, [n_local]
    counts = [None] * ws
    dist.all_gather_object(counts, int(n_local), group=group)
    labels = [None] * ws
    dist.all_gather_object(labels, y_local, group=group)
    return int(sum(counts[:rank])), np.concatenate(labels), counts


def gather_pairs(d2: torch.Tensor, idx: torch.Tensor, group=None):
    """All-gather per-rank (distance, global index) tensors ([nq] or [nq, k]) -> ([P, ...] float32, [P, ...] int64) on
    d2's device."""
    rank, ws = world(group)
    if ws == 1:
        return d2.unsqueeze(0), idx.unsqueeze(0)
    ds = [torch.empty_like(d2) for _ in range(ws)]
    js = [torch.empty_like(idx) for _ in range(ws)]
    dist.all_gather(ds, d2.contiguous(), group=group)
    dist.all_gather(js, idx.contiguous(), group=group)
    return torch.stack(ds).contiguous(), torch.stack(js).contiguous()


def merge_topk(d_all: torch.Tensor, i_all: torch.Tensor, k: int):
    """Per-shard k-NN lists gathered as [P, nq, k] -> the global k nearest per query, ascending by (distance, index):
    [nq, k] float32 / int64.  Shards that returned fewer than k rows pad with (inf, -1), which sort last."""
    P, nq, kk = d_all.shape
    d = d_all.permute(1, 0, 2).reshape(nq, P * kk)
    i = i_all.permute(1, 0, 2).reshape(nq, P * kk)
    big = torch.iinfo(torch.int64).max
    order = torch.argsort(torch.where(i < 0, torch.full_like(i, big), i), dim=1, stable=True)   # ties -> lowest index
    d, i = torch.gather(d, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(d, dim=1, stable=True)[:, :k]
    return torch.gather(d, 1, order).contiguous(), torch.gather(i, 1, order).contiguous()


def gather_rows(x: torch.Tensor, group=None):
    """Concatenate per-rank row blocks (possibly of different lengths) in rank order, e.g. (B_r, D) embeddings."""
    rank, ws = world(group)
    if ws == 1:
        return x
    counts = [None] * ws
    dist.all_gather_object(counts, int(x.shape[0]), group=group)
    m = max(counts)
    pad = torch.zeros((m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])
