"""Run under torchrun on G GPUs: data-parallel train step and catalog-sharded eval must equal the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from tcar_b200 import parallel, synth  # noqa: E402
from tcar_b200.model_combine import Seq2SeqAttNN  # noqa: E402


def build(N, rank, world, **extra):
    content, mwdhm, _ = synth.make_catalog(N, seed=3)
    np.random.seed(2020)
    args = dict(extra, publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={}, reverse_item=None,
                content_emb=content, emb_stddev=0.04, stddev=0.05, hidden_size=250, time_hidden_size=64, l2_emb=0.0,
                batch_size=512, epoch=1, neg_num=20, lr=0.001, max_grad=150, rank=rank, world_size=world)
    return Seq2SeqAttNN(args), mwdhm


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, B, T, Nn = 20000, 301, 5, 20
    res = {}
    # ---- data-parallel train step vs the same global batch on one GPU
    model, mwdhm = build(N, rank, world)
    packed = synth.make_index_batch(N, B, T, Nn, mwdhm, seed=5)
    pl, Bl, _, _ = parallel.shard_packed(packed, B, T, Nn, rank, world)
    bt = model.to_device(torch.from_numpy(pl).pin_memory(), Bl, T, Nn)
    loss_local = model.train_step(bt).clone()
    torch.cuda.synchronize()
    g_theta, item_after = model.ps.theta_g.clone(), model.ps.item.clone()
    if model._sharded_update():
        # reduce-scatter path: every rank holds the reduced gradient of its own row slice only
        per = model.ps.rows_alloc // world
        g_item, g_rows = model._g_slice.clone(), slice(rank * per, (rank + 1) * per)
    else:
        g_item, g_rows = model.ps.item_g_full.clone(), slice(0, model.ps.rows_alloc)
    single, _ = build(N, 0, 1)
    btf = single.to_device(torch.from_numpy(packed).pin_memory(), B, T, Nn)
    loss_full = single.train_step(btf).clone()
    torch.cuda.synchronize()
    lo, hi = parallel.shard_sessions(B, rank, world)
    rel = lambda a, b: float((a - b).double().norm() / (b.double().norm() + 1e-30))
    res["loss_maxabs"] = float((loss_local - loss_full[lo:hi]).abs().max())
    res["g_item_rel"] = rel(g_item, single.ps.item_g_full[g_rows])
    # gradients differ in the last bits (different summation order), so a few bf16 roundings of the refreshed scoring
    # operand may flip: require them to be rare and one ulp at most
    diff = (model.ps.iext.float() - single.ps.iext.float()).abs()
    res["iext_mismatch_frac"] = float((diff > 0).float().mean())
    res["iext_equal"] = bool(res["iext_mismatch_frac"] < 1e-3 and float(diff.max()) <= 2 ** -7 * float(single.ps.iext.float().abs().max()))
    res["g_theta_rel"] = rel(g_theta, single.ps.theta_g)
    res["item_after_rel"] = rel(item_after, single.ps.item)
    # checkpoint under the sharded data-parallel update: every rank owns the Adam moments of its row slice only;
    # save_model must gather them (util.save_model -> sync_optimizer_state) so that the file equals a 1-GPU checkpoint
    import tempfile
    from tcar_b200 import util
    with tempfile.TemporaryDirectory() as d:
        path = util.save_model(model, dict(dataset="synth/", split_way="Normal/", foldnum=0, modelpath=d + "/r%d/" % rank))
        dist.barrier()
        if rank == 0:
            sd = torch.load(path, map_location="cpu", weights_only=True)
            want_m, want_v = single.ps.item_m[:, :250].cpu(), single.ps.item_v[:, :250].cpu()
            res["ckpt_moments_rel"] = max(rel(sd["item_m"], want_m), rel(sd["item_v"], want_v))
            res["ckpt_nonzero_rows"] = int((sd["item_v"].abs().sum(1) > 0).sum())
        else:
            res["ckpt_moments_rel"] = 0.0
        dist.barrier()
    # ---- catalog-sharded eval vs single GPU
    epacked = synth.make_index_batch(N, B, T, 0, mwdhm, seed=8)
    ebt = single.to_device(torch.from_numpy(epacked).pin_memory(), B, T, 0)
    top1, n1, ce1 = [x.clone() for x in single.eval_step(ebt)]
    single.world, single.rank = world, rank
    slo, shi = single.shard_bounds(world)[rank]
    topg, ng, ceg = single.eval_step(ebt, shard=(slo, shi, single.iext_shard(slo, shi)))
    torch.cuda.synchronize()
    res["top20_equal"] = bool(torch.equal(top1, topg))
    hit = n1 < 20
    res["rank_equal"] = bool(torch.equal(hit, ng < 20) and torch.equal(n1[hit], ng[hit]))
    res["ce_maxabs"] = float((ce1 - ceg).abs().max())
    # ---- catalog-sharded train steps (catalog_parallel.py) vs the same global batches on one GPU.  Two steps: the
    # second one gathers rows that OTHER ranks updated in the first (peer loads), and B2 < world exercises empty slices
    cat, _ = build(N, rank, world, train_parallel="catalog")
    ref, _ = build(N, 0, 1)
    rel = lambda a, b: float((a - b).double().norm() / (b.double().norm() + 1e-30))
    cres = {"loss_maxabs": 0.0, "g_item_rel": 0.0, "g_theta_rel": 0.0}
    rb = cat._cat_row_bounds
    own = slice(rb[rank], rb[rank + 1])
    plan = []
    for step, (Bs, Ts) in enumerate([(B, T), (world - 1, 3), (512, 2)]):
        pk = synth.make_index_batch(N, Bs, Ts, Nn, mwdhm, seed=50 + step)
        pl, Bl, _, _ = parallel.shard_packed(pk, Bs, Ts, Nn, rank, world)
        cbt = cat.to_device(torch.from_numpy(pl).pin_memory(), Bl, Ts, Nn)
        cbt.counts = parallel.catalog_counts(Bs, world)
        plan.append((Bs, Ts, pk, cbt))
    cres["lookahead"] = bool(cat.cat_lookahead)        # TCAR_CATALOG_LOOKAHEAD=1: next batch passed to every step
    for step, (Bs, Ts, pk, cbt) in enumerate(plan):
        nxt = plan[step + 1][3] if (cat.cat_lookahead and step + 1 < len(plan)) else None
        lc = cat.train_step(cbt, nxt).clone()
        lr_ = ref.train_step(ref.to_device(torch.from_numpy(pk).pin_memory(), Bs, Ts, Nn)).clone()
        torch.cuda.synchronize()
        lo, hi = parallel.shard_sessions(Bs, rank, world)
        if hi > lo:
            cres["loss_maxabs"] = max(cres["loss_maxabs"], float((lc - lr_[lo:hi]).abs().max()))
        cres["g_item_rel"] = max(cres["g_item_rel"], rel(cat.ps.item_g_full[own], ref.ps.item_g_full[own]))
        cres["g_theta_rel"] = max(cres["g_theta_rel"], rel(cat.ps.theta_g, ref.ps.theta_g))
        # clip norm of the item gradient: fused (GEMM partials + scatter corrections of every group), summed over ranks
        want = float((ref.ps.item_g.double() ** 2).sum())
        cres["item_sqnorm_rel"] = max(cres.get("item_sqnorm_rel", 0.0), abs(float(cat._sq_slot.item()) - want) / want)
    cres["own_rows_rel"] = rel(cat.ps.item_full[own], ref.ps.item_full[own])
    cat.sync_item_table()
    torch.cuda.synchronize()
    cres["item_after_rel"] = rel(cat.ps.item, ref.ps.item)
    cres["theta_after_rel"] = rel(cat.ps.theta, ref.ps.theta)
    cres["moments_rel"] = max(rel(cat.ps.item_m, ref.ps.item_m), rel(cat.ps.item_v, ref.ps.item_v))
    cdiff = (cat.ps.iext.float() - ref.ps.iext.float()).abs()
    cres["iext_mismatch_frac"] = float((cdiff > 0).float().mean())
    # three consecutive steps: from the second one on the two runs start from parameters that already differ in the
    # last bits (Adam normalises every element by its own |g|), so the bounds are looser than for a single step
    cres["ok"] = bool(cres["loss_maxabs"] < 1e-4 and cres["g_item_rel"] < 5e-5 and cres["g_theta_rel"] < 5e-5
                      and cres["own_rows_rel"] < 5e-6 and cres["item_after_rel"] < 5e-6
                      and cres["theta_after_rel"] < 5e-6 and cres["moments_rel"] < 5e-5 and cres["item_sqnorm_rel"] < 1e-4
                      and cres["iext_mismatch_frac"] < 1e-3)
    res["catalog"] = cres
    cat.close_peers()
    ok = (cres["ok"] and res["loss_maxabs"] < 1e-5 and res["g_item_rel"] < 1e-5 and res["g_theta_rel"] < 1e-5
          and res["item_after_rel"] < 1e-6 and res["iext_equal"] and res["top20_equal"] and res["rank_equal"]
          and res["ce_maxabs"] < 1e-4 and res["ckpt_moments_rel"] < 1e-5)
    res["ok"] = ok
    print(f"rank {rank}/{world}: " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
