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


# One shard's evaluation results for <= 512 queries as ONE block of 32-bit words (TCAR_EVAL_OFF_* in
# include/tcar_b200.h): [scores 512x20 f32 | ids 512x20 i32 | n_greater 512 i32 | sumexp 512 f32 | rowmax 512 f32].
QROWS = 512
EVAL_OFF_SCORES, EVAL_OFF_IDS, EVAL_OFF_NGT = 0, QROWS * TOPK, 2 * QROWS * TOPK
EVAL_OFF_SUMEXP, EVAL_OFF_ROWMAX, EVAL_BLOCK_WORDS = EVAL_OFF_NGT + QROWS, EVAL_OFF_NGT + 2 * QROWS, EVAL_OFF_NGT + 3 * QROWS


def pack_eval_block(top_ids, top_scores, n_greater, sumexp, rowmax=None):
    """Build a result block from separate tensors (the CUDA path writes the planes in place; CPU tests use this)."""
    B = top_ids.shape[0]
    blk = torch.zeros(EVAL_BLOCK_WORDS, dtype=torch.float32, device=top_ids.device)
    bi = blk.view(torch.int32)
    blk[EVAL_OFF_SCORES: EVAL_OFF_SCORES + B * TOPK] = top_scores.reshape(-1).float()
    bi[EVAL_OFF_IDS: EVAL_OFF_IDS + B * TOPK] = top_ids.reshape(-1).int()
    bi[EVAL_OFF_NGT: EVAL_OFF_NGT + B] = n_greater.int()
    blk[EVAL_OFF_SUMEXP: EVAL_OFF_SUMEXP + B] = sumexp.float()
    if rowmax is not None:
        blk[EVAL_OFF_ROWMAX: EVAL_OFF_ROWMAX + B] = rowmax.float()
    return blk


def gather_eval_blocks(block, world, out=None):
    """THE exchange of the catalog-sharded evaluation: one all-gather of every rank's result block -> [world, WORDS]
    (rank-major).  The merge (tcar_eval_merge in the product path) then runs locally on every rank."""
    import torch.distributed as dist
    if out is None:
        out = torch.empty(world, EVAL_BLOCK_WORDS, dtype=torch.float32, device=block.device)
    dist.all_gather_into_tensor(out.view(-1), block.contiguous())
    return out
