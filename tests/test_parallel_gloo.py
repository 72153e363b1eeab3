"""Host-side multi-GPU logic on CPU: world_size 2 over gloo.  The CUDA merge kernel is replaced by the
oracle's merge here -- the same callable slot the product fills with tcar_topk_merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tcar_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import __graft_entry__ as ge
    ge.load_package()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


# ---------------------------------------------------------------------------------------------- pure host logic
@pytest.mark.parametrize("B,world", [(512, 2), (7, 4), (3, 8), (1, 2), (512, 8)])
def test_shard_sessions_partition(B, world):
    from tcar_b200 import parallel
    spans = [parallel.shard_sessions(B, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == B
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("B,T,Nn,world", [(10, 3, 4, 2), (5, 1, 0, 4), (64, 20, 20, 8)])
def test_shard_packed_reassembles(B, T, Nn, world):
    from tcar_b200 import parallel, synth
    N = 500
    _, mwdhm, _ = synth.make_catalog(N, seed=1)
    packed = synth.make_index_batch(N, B, T, Nn, mwdhm, seed=3)
    whole = synth.unpack(packed, B, T, Nn)
    parts = []
    for r in range(world):
        pl, Bl, Tl, Nl = parallel.shard_packed(packed, B, T, Nn, r, world)
        assert pl.dtype == np.int32 and pl.size == 7 * Bl * T + 3 * Bl + Bl * Nn
        if Bl:
            parts.append(synth.unpack(pl, Bl, Tl, Nl))
    for k in whole:
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts]), whole[k])


@pytest.mark.parametrize("N,G", [(364047, 8), (364047, 2), (1000, 8), (300, 4)])
def test_shard_bounds_cover_catalog(N, G):
    from tcar_b200 import parallel
    n_pad = (N + 255) // 256 * 256
    b = parallel.shard_bounds(N, n_pad, G)
    assert b[0][0] == 0 and max(hi for _, hi in b) == N
    covered = sum(hi - lo for lo, hi in b)
    assert covered == N
    assert all(lo % 256 == 0 for lo, hi in b if hi > lo)


# ---------------------------------------------------------------------------------------------- world_size 2, gloo
def _dp_allreduce(rank, world):
    from tcar_b200 import parallel
    g = torch.full((5, 3), float(rank + 1))
    h = torch.arange(4, dtype=torch.float32) * (rank + 1)
    parallel.allreduce_sum((g, h), world)
    return g.numpy().copy(), h.numpy().copy()


def test_gradient_allreduce_is_a_sum():
    out = _spawn(_dp_allreduce, 2)
    for g, h in out:
        np.testing.assert_array_equal(g, np.full((5, 3), 3.0))
        np.testing.assert_array_equal(h, np.arange(4) * 3.0)


def _sharded_eval(rank, world):
    """Every rank scores all queries against its catalog shard (numpy stand-in for the kernels), then the product's
    gather/merge plumbing must reproduce the unsharded top-20, rank counts and softmax denominators."""
    from tcar_b200 import parallel
    rs = np.random.RandomState(0)
    B, N = 37, 3000
    S = rs.normal(size=(B, N)).astype(np.float32)
    S[:, 100:140] = S[:, 200:240]                    # exact ties across and inside shards
    label = rs.randint(0, N, B)
    lo, hi = parallel.shard_bounds(N, (N + 255) // 256 * 256, world)[rank]
    loc = S[:, lo:hi]
    ids = np.arange(lo, hi)
    top_i = np.full((B, 20), -1, np.int32)
    top_s = np.full((B, 20), -np.inf, np.float32)
    for b in range(B):
        o = np.lexsort((ids, -loc[b]))[:20]
        top_i[b, : len(o)] = ids[o]
        top_s[b, : len(o)] = loc[b, o]
    ngt = torch.from_numpy((loc > S[np.arange(B), label][:, None]).sum(1).astype(np.int32))
    # softmax partial sums relative to the label score; the second rank's are additionally shifted by its row maximum
    # (what the overflow guard does for a shard whose best score beats the label by more than the limit)
    c = S[np.arange(B), label].astype(np.float64)
    arg2 = (loc.astype(np.float64) - c[:, None]) * np.log2(np.e)
    rowmax = arg2.max(1)
    if rank == 1:
        rowmax += 100.0                               # pretend huge margins: forces the shifted representation
        arg2 += 100.0
    shift = np.where(rowmax > O.EXP_LIMIT2, rowmax, 0.0)
    sumexp = torch.from_numpy(np.exp2(arg2 - shift[:, None]).sum(1).astype(np.float32))
    blk = parallel.pack_eval_block(torch.from_numpy(top_i), torch.from_numpy(top_s), ngt, sumexp,
                                   torch.from_numpy(rowmax.astype(np.float32)))
    blocks = parallel.gather_eval_blocks(blk, world)                  # ONE collective
    oi, os_, ngt, ce = O.merge_eval_blocks(blocks.numpy(), B)
    return oi, ngt, ce, S, label, lo, hi


def test_catalog_sharded_eval_equals_unsharded():
    out = _spawn(_sharded_eval, 2)
    for oi, ngt, ce, S, label, lo, hi in out:
        np.testing.assert_array_equal(oi, O.top20(S))
        ref_rank = np.array([int((row[l] < row).sum()) for row, l in zip(S, label)])
        np.testing.assert_array_equal(ngt, ref_rank)
        # reference CE with the second shard's scores raised by 100 log2 units (as the worker pretended)
        S2 = S.astype(np.float64).copy()
        lo1, hi1 = out[1][5], out[1][6]
        S2[:, lo1:hi1] += 100.0 * np.log(2.0)
        c = S[np.arange(len(label)), label].astype(np.float64)
        m = S2.max(1)
        want = np.log(np.exp(S2 - m[:, None]).sum(1)) + m - c
        np.testing.assert_allclose(ce, want, rtol=1e-5)


# ---------------------------------------------------------------------------------------------- catalog-sharded training
@pytest.mark.parametrize("N,G", [(364047, 8), (364047, 2), (700, 8), (300, 4), (256, 2)])
def test_catalog_row_bounds_cover_table(N, G):
    """Table-row ownership of the catalog-sharded train step: every one of the N + 1 rows has exactly one owner, the
    pad row 0 goes with the first shard, and a shard's rows are its item ids + 1."""
    from tcar_b200 import parallel
    n_pad = (N + 255) // 256 * 256
    b = parallel.shard_bounds(N, n_pad, G)
    rb = parallel.catalog_row_bounds(b, N)
    assert len(rb) == G + 1 and rb[0] == 0 and rb[-1] == N + 1
    assert all(x <= y for x, y in zip(rb, rb[1:]))
    for g, (lo, hi) in enumerate(b):
        if hi > lo:
            assert rb[g + 1] == hi + 1 and (rb[g] == lo + 1 or (g == 0 and rb[g] == 0))
    with pytest.raises(ValueError):
        parallel.catalog_row_bounds([(0, 10), (10, 20)], 30)


@pytest.mark.parametrize("B,world", [(512, 8), (7, 4), (1, 2), (0, 2)])
def test_catalog_counts_match_shard_packed(B, world):
    from tcar_b200 import parallel
    c = parallel.catalog_counts(B, world)
    assert sum(c) == B and len(c) == world
    assert c == [hi - lo for lo, hi in (parallel.shard_sessions(B, r, world) for r in range(world))]


def _catalog_softmax_worker(rank, world):
    """The exchange pattern of catalog_parallel.train_step_catalog with CPU tensors (fp64): all-gather the sessions'
    query rows, score against the OWN item range, all-reduce the softmax partial sums, sum the dQ partials over ranks
    and keep the own sessions' rows, all-gather Qs, local dItems.  Compared with autograd on the full problem."""
    from tcar_b200 import parallel
    torch.manual_seed(0)
    N, K, Bg = 700, 24, 5
    n_pad = (N + 255) // 256 * 256
    items = torch.randn(N, K, dtype=torch.float64)
    Qall = torch.randn(world * Bg, K, dtype=torch.float64)
    labels = torch.randint(0, N, (world * Bg,))
    # full problem (every rank computes the same reference)
    Qr, Ir = Qall.clone().requires_grad_(True), items.clone().requires_grad_(True)
    S = Qr @ Ir.t()
    ce = torch.logsumexp(S, 1) - S.gather(1, labels[:, None]).squeeze(1)
    ce.sum().backward()
    # sharded
    lo, hi = parallel.shard_bounds(N, n_pad, world)[rank]
    mine = slice(rank * Bg, (rank + 1) * Bg)
    q_loc = Qall[mine].contiguous()
    c_loc = (q_loc * items[labels[mine]]).sum(1)                       # label score: the fetched label rows
    q_all = torch.empty(world * Bg, K, dtype=torch.float64)
    c_all = torch.empty(world * Bg, dtype=torch.float64)
    dist.all_gather_into_tensor(q_all, q_loc)
    dist.all_gather_into_tensor(c_all, c_loc)
    E = torch.exp(q_all @ items[lo:hi].t() - c_all[:, None])           # [R*Bg, n_loc]
    sumexp = E.sum(1)
    dist.all_reduce(sumexp)
    dq = E @ items[lo:hi]                                              # partial over the own items
    dist.all_reduce(dq)                                                # (reduce-scatter in the product)
    dq_loc = dq[mine] / sumexp[mine, None] - items[labels[mine]]
    qs_all = torch.empty(world * Bg, K, dtype=torch.float64)
    dist.all_gather_into_tensor(qs_all, (q_loc / sumexp[mine, None]).contiguous())
    d_items = E.t() @ qs_all                                           # dense rows of the own range, complete
    # sparse label rows of EVERY rank's sessions that fall into the own range (ranged scatter)
    for b in range(world * Bg):
        n = int(labels[b])
        if lo <= n < hi:
            d_items[n - lo] -= q_all[b]
    ok_ce = torch.allclose(torch.log(sumexp[mine]), ce[mine].detach(), rtol=1e-10, atol=1e-10)
    ok_dq = torch.allclose(dq_loc, Qr.grad[mine], rtol=1e-9, atol=1e-11)
    ok_di = torch.allclose(d_items, Ir.grad[lo:hi], rtol=1e-9, atol=1e-11)
    return bool(ok_ce and ok_dq and ok_di)


def test_catalog_sharded_softmax_exchange_is_exact():
    assert all(_spawn(_catalog_softmax_worker, world=2))
