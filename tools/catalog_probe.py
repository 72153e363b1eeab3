"""Phase breakdown of the catalog-sharded train step (catalog_parallel.py) under torchrun: GPU time between the phase
marks of train_step_catalog (CUDA events on the main stream), host enqueue time per step, GPU time per step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/catalog_probe.py
"""
import json
import os
import sys
import time

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
    N, B, Nn = int(os.environ.get("PROBE_ITEMS", 364047)), 512, 20
    Ts = [int(t) for t in os.environ.get("PROBE_T", "8,4,2,1,1,1,1,1,4,1,1,3,3,1,1,2").split(",")]
    content, mwdhm, _ = synth.make_catalog(N)
    np.random.seed(2020)
    args = dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={}, reverse_item=None,
                content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250, time_hidden_size=64, l2_emb=0.0,
                batch_size=B, epoch=1, neg_num=Nn, lr=0.001, max_grad=150, rank=rank, world_size=world,
                train_parallel="catalog")
    model = Seq2SeqAttNN(args)
    dev = [model.to_device(torch.from_numpy(synth.make_index_batch(N, B, t, Nn, mwdhm, seed=1000 * (rank + 1) + i))
                           .pin_memory(), B, t, Nn) for i, t in enumerate(Ts)]
    for i in range(6):
        model.train_step(dev[i % len(dev)])
    torch.cuda.synchronize()
    dist.barrier()
    K = 32
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(K):
        model.train_step(dev[i % len(dev)])
    e1.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    gpu_ms = e0.elapsed_time(e1) / K
    dist.barrier()
    phases = {}
    for i in range(K):
        model._cat_trace = []
        model.train_step(dev[i % len(dev)])
        torch.cuda.synchronize()
        tr = model._cat_trace
        for (_, a), (name, b) in zip(tr, tr[1:]):
            phases[name] = phases.get(name, 0.0) + a.elapsed_time(b) * 1000 / K
    model._cat_trace = None
    out = {"rank": rank, "world": world, "gpu_ms_per_step": gpu_ms, "host_enqueue_ms_per_step": t_enq * 1e3 / K,
           "wall_ms_per_step": t_all * 1e3 / K, "phase_us": {k: round(v, 1) for k, v in phases.items()},
           "phase_sum_us": round(sum(phases.values()), 1)}
    print("probe " + json.dumps(out), flush=True)
    dist.barrier()
    model.close_peers()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
