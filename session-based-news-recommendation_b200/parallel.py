"""Multi-GPU plumbing of the TCAR hot path (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in the
CPU tests).  The reference has no distributed code at all ("Parallelism strategies: none"); the two
partitionings below are the ones BASELINE.json's north_star prescribes:

  training    data parallel over the sessions of a length bucket, ONE all-reduce(SUM) of the gradients per step.
              SUM, not mean: the reference differentiates the batch SUM of the [B,1] loss (model_combine.py:156), and
              the per-tensor clip_by_norm (:158-160) must see the gradient of the whole global batch.
  training    (catalog_parallel.py) alternatively the CATALOG is split for training too: every rank
              owns the parameters, moments and gradient of a contiguous item range and scores all sessions against it;
              only session-sized tensors are exchanged.
  evaluation  the item catalog is split into contiguous id ranges; every rank scores all B queries against its own
              range, then all-gather of the per-rank top-20 (score, id) lists + merge by (score desc, id asc), and
              all-reduce(SUM) of the rank counts #(S > S[label]) and of the softmax partial sums.

Everything here is device-agnostic host logic: tensors stay wherever the caller put them and the merge itself is a
callable (the CUDA kernel tcar_topk_merge in the product path).
"""
import numpy as np
import torch

TOPK = 20


def is_distributed(world):
    return world > 1 and torch.distributed.is_available() and torch.distributed.is_initialized()


# ------------------------------------------------------------------------------------------------- training (DP)
def shard_sessions(B, rank, world):
    """[lo, hi) of the B sessions of one batch owned by `rank`: contiguous, balanced, covers every session once."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_packed(packed, B, T, Nn, rank, world):
    """Slice one packed int32 batch [7*B*T idx | 2*B ctx | B label | B*Nn neg] (model_combine.Batch) down to the
    sessions owned by `rank`.  Returns (packed_local, B_local, T, Nn); B_local may be 0 for tiny tail batches."""
    lo, hi = shard_sessions(B, rank, world)
    M = B * T
    packed = np.asarray(packed)
    idx = packed[: 7 * M].reshape(7, B, T)[:, lo:hi]
    ctx = packed[7 * M: 7 * M + 2 * B].reshape(2, B)[:, lo:hi]
    label = packed[7 * M + 2 * B: 7 * M + 3 * B][lo:hi]
    parts = [idx.reshape(-1), ctx.reshape(-1), label]
    if Nn:
        parts.append(packed[7 * M + 3 * B:].reshape(B, Nn)[lo:hi].reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts).astype(np.int32)), hi - lo, T, Nn


def allreduce_sum(tensors, world):
    """Gradient all-reduce of the data-parallel train step (in place, SUM)."""
    if not is_distributed(world):
        return
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


# ------------------------------------------------------------------------------------------------- evaluation
def shard_bounds(N, n_pad, G, align=256):
    """Contiguous item-id ranges [lo, hi) per rank, aligned to `align` rows so that a shard of the bf16 scoring
    operand is a plain row-slice of the full one.  Ranks past the end of a small catalog get an empty range."""
    tiles = n_pad // align
    per = (tiles + G - 1) // G
    return [(min(g * per * align, N), min((g + 1) * per * align, N)) for g in range(G)]


def catalog_row_bounds(bounds, N):
    """Item-TABLE row ranges of a catalog-sharded train step (catalog_parallel.py) from the 0-based item-id ranges of
    shard_bounds(): table row = item id + 1, and the first shard also owns the pad row 0.  Returns [G + 1] row
    boundaries: shard g owns rows [rb[g], rb[g + 1]); together they cover the N + 1 rows exactly once."""
    rb = [0] + [hi + 1 for _, hi in bounds]
    if rb[-1] != N + 1 or any(b < a for a, b in zip(rb, rb[1:])):
        raise ValueError("shard bounds must be contiguous, ascending and end at N")
    return rb


def catalog_counts(B, world):
    """Sessions per rank when one global batch of B sessions is split with shard_sessions (every rank computes the
    same list without communicating)."""
    return [hi - lo for lo, hi in (shard_sessions(B, g, world) for g in range(world))]


def gather_merge_topk(top_ids, top_scores, n_greater, sumexp, world, merge):
    """All-gather the per-rank top-20 lists and merge them; sum the rank counts and the softmax partial sums.
    `merge(ids [G,B,20] int32, scores [G,B,20] f32) -> (ids [B,20], scores [B,20])` must order by
    (score desc, id asc) and treat id < 0 as an empty slot."""
    import torch.distributed as dist
    B = top_ids.shape[0]
    ids = torch.empty(world * B, TOPK, device=top_ids.device, dtype=torch.int32)
    sc = torch.empty(world * B, TOPK, device=top_ids.device, dtype=torch.float32)
    dist.all_gather_into_tensor(ids, top_ids.contiguous())          # rank-major concatenation along dim 0
    dist.all_gather_into_tensor(sc, top_scores.contiguous())
    ids, sc = ids.view(world, B, TOPK), sc.view(world, B, TOPK)
    dist.all_reduce(n_greater, op=dist.ReduceOp.SUM)
    dist.all_reduce(sumexp, op=dist.ReduceOp.SUM)
    out_ids, out_sc = merge(ids, sc)
    return out_ids, out_sc, n_greater, sumexp
