"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the hot path shards.

  * embedding extraction: images are independent -> split the batch, no data-path collective; `gather_rows` collects the
    (B, D) embeddings when one rank needs them all.
  * 1-NN: gallery row-sharded in contiguous blocks; every rank scores all queries against its shard, the per-rank
    (squared distance, global index) pairs are all-gathered (8-12 bytes per query per rank) and merged with
    lowest-index tie-break by hfr_knn_merge (packed {fp64 dist2, int64 index} records, one all_gather_into_tensor).

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


def gather_neighbors(rec: torch.Tensor, group=None):
    """All-gather the packed neighbour records of every rank ([nq, k, 2] int64 = hfr_neighbor {double dist2; int64
    index}) into [P, nq, k, 2] with ONE collective; the merge is hfr_knn_merge."""
    rank, ws = world(group)
    if ws == 1:
        return rec.unsqueeze(0)
    out = torch.empty((ws,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
    if rec.is_cuda:
        dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    else:                           # gloo (CPU tests): the list form, written straight into the slices of `out`
        dist.all_gather(list(out.unbind(0)), rec.contiguous(), group=group)
    return out


def gather_rows(x: torch.Tensor, group=None, total=None):
    """Concatenate per-rank row blocks (possibly of different lengths) in rank order, e.g. (B_r, D) embeddings.
    total: number of rows over all ranks when the blocks follow shard_rows(total, world, rank) - then no metadata is
    exchanged (one all-gather, no host synchronisation); otherwise the counts travel first."""
    rank, ws = world(group)
    if ws == 1:
        return x
    if total is not None:
        counts = [shard_rows(total, ws, r)[1] - shard_rows(total, ws, r)[0] for r in range(ws)]
        if counts[rank] != x.shape[0]:
            raise ValueError(f"rank {rank} holds {x.shape[0]} rows, shard_rows({total}, {ws}, {rank}) says {counts[rank]}")
    else:
        counts = [None] * ws
        dist.all_gather_object(counts, int(x.shape[0]), group=group)
    m = max(counts)
    if min(counts) == m:
        pad = x.contiguous()
    else:
        pad = torch.zeros((m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[: x.shape[0]] = x
    out = torch.empty((ws, m) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if x.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), pad, group=group)
    if min(counts) == m:
        return out.reshape((ws * m,) + tuple(x.shape[1:]))
    return torch.cat([out[r, :c] for r, c in enumerate(counts)])
