"""Per-phase GPU time of Seq2SeqAttNN.eval_round under torchrun (max over ranks), and the round throughput.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/eval_round_probe.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from tcar_b200 import synth  # noqa: E402
from tcar_b200.model_combine import Seq2SeqAttNN  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, B = int(os.environ.get("ITEMS", "364047")), 512
    content, mwdhm, _ = synth.make_catalog(N)
    np.random.seed(2020)
    model = Seq2SeqAttNN(dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={},
                              reverse_item=None, content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250,
                              time_hidden_size=64, l2_emb=0.0, batch_size=B, epoch=1, neg_num=20, lr=0.001, max_grad=150,
                              rank=rank, world_size=world))
    rs = np.random.RandomState(2020)
    pr = 0.55 ** np.arange(1, 21)
    Ts = [int(t) for t in rs.choice(np.arange(1, 21), size=16, p=pr / pr.sum())]
    bts = [model.to_device(torch.from_numpy(synth.make_index_batch(N, B, t, 0, mwdhm, seed=77 + i + 100 * rank)).pin_memory(),
                           B, t, 0) for i, t in enumerate(Ts)]
    counts = [B] * world
    for two_stage in (True, False):
        for i in range(5):
            model.eval_round(bts[i % 16], counts, two_stage=two_stage)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            model.eval_round(bts[i % 16], counts, two_stage=two_stage)
        e1.record()
        if two_stage:
            torch.cuda.synchronize()
            dist.barrier()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record()
            for i in range(20):
                model.eval_round(bts[i % 16], counts, two_stage=True, next_bt=bts[(i + 1) % 16])
            model.sync_updates()
            e3.record()
            torch.cuda.synchronize()
            t2 = torch.tensor([e2.elapsed_time(e3) / 20], device=model.dev, dtype=torch.float64)
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"two_stage=True + look-ahead: {float(t2) * 1e3:.1f} us/round, {world * B / (float(t2) * 1e-3):.0f} queries/s", flush=True)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=model.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"two_stage={two_stage}: {float(t) * 1e3:.1f} us/round, {world * B / (float(t) * 1e-3):.0f} queries/s", flush=True)
    acc = {}
    for rep in range(6):
        model._round_trace = []
        torch.cuda.synchronize()
        dist.barrier()
        model.eval_round(bts[rep % 16], counts, two_stage=True)
        torch.cuda.synchronize()
        tr = model._round_trace
        for (n0, a), (n1, b) in zip(tr, tr[1:]):
            if rep >= 2:
                acc.setdefault(n1, []).append(a.elapsed_time(b) * 1e3)
    model._round_trace = None
    names = list(acc)
    t = torch.tensor([np.mean(acc[n]) for n in names], device=model.dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("two-stage round, per phase (us, max over ranks, round synchronised at the start):")
        for n, v in zip(names, t.tolist()):
            print(f"  {n:18s} {v:8.1f}")
        print(f"  {'total':18s} {sum(t.tolist()):8.1f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
