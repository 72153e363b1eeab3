"""Where the end-to-end train loop (pinned host batch -> H2D -> train_step -> D2H of the loss) spends its time on ONE
GPU: the bench's e2e loop in variants, each with GPU ms/step (CUDA events) and the host's time per piece.

    python tools/e2e_probe.py            (PROBE_ITEMS, PROBE_T as in tools/catalog_probe.py)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from tcar_b200 import synth  # noqa: E402
from tcar_b200.model_combine import Seq2SeqAttNN  # noqa: E402


def main():
    torch.cuda.set_device(0)
    N, B, Nn = int(os.environ.get("PROBE_ITEMS", 364047)), 512, 20
    Ts = [int(t) for t in os.environ.get("PROBE_T", "8,4,2,1,1,1,1,1,4,1,1,3,3,1,1,2").split(",")]
    K = int(os.environ.get("PROBE_STEPS", 48))
    content, mwdhm, _ = synth.make_catalog(N)
    np.random.seed(2020)
    args = dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={}, reverse_item=None,
                content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250, time_hidden_size=64, l2_emb=0.0,
                batch_size=B, epoch=1, neg_num=Nn, lr=0.001, max_grad=150, rank=0, world_size=1, train_parallel="dp")
    model = Seq2SeqAttNN(args)
    host = [torch.from_numpy(synth.make_index_batch(N, B, t, Nn, mwdhm, seed=1000 + i)).pin_memory()
            for i, t in enumerate(Ts)]
    dev = [model.to_device(h, B, t, Nn) for h, t in zip(host, Ts)]
    n = len(Ts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run(name, h2d, fetch, lag, look=True, steps=K):
        """h2d: copy the next batch from pinned memory every step; fetch: D2H of the loss every step; lag: how many
        steps later the host reads a loss (0 = at once, as a blocking .cpu())."""
        for rep in range(2):                      # first repetition = warm-up
            t = {"h2d": 0.0, "step": 0.0, "fetch": 0.0, "wait": 0.0}
            bt = model.to_device(host[0], B, Ts[0], Nn) if h2d else dev[0]
            pend = []
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            e0.record()
            for i in range(steps):
                j = (i + 1) % n
                a = time.perf_counter()
                nb = model.to_device(host[j], B, Ts[j], Nn) if h2d else dev[j]
                b = time.perf_counter()
                loss = model.train_step(bt, nb if look else None)
                c = time.perf_counter()
                if fetch:
                    pend.append(model.fetch_async(loss))
                d = time.perf_counter()
                while len(pend) > lag:
                    float(pend.pop(0).get().sum())
                e = time.perf_counter()
                t["h2d"] += b - a
                t["step"] += c - b
                t["fetch"] += d - c
                t["wait"] += e - d
                bt = nb
            model.sync_updates()
            for h in pend:
                h.get()
            e1.record()
            enq = time.perf_counter() - w0
            torch.cuda.synchronize()
        out = {"variant": name, "gpu_ms_per_step": round(e0.elapsed_time(e1) / steps, 4),
               "host_ms_per_step": round(enq * 1e3 / steps, 4),
               "host_us": {k: round(v * 1e6 / steps, 1) for k, v in t.items()}}
        print("e2e_probe " + json.dumps(out), flush=True)

    run("resident, no fetch", False, False, 0)
    run("resident, fetch lag 1", False, True, 1)
    run("resident, fetch lag 2", False, True, 2)
    run("h2d, no fetch", True, False, 0)
    run("h2d, fetch lag 0 (blocking)", True, True, 0)
    run("h2d, fetch lag 1 (bench e2e)", True, True, 1)
    run("h2d, fetch lag 2", True, True, 2)
    run("h2d, fetch lag 1, no look-ahead", True, True, 1, look=False)
    run("resident, no fetch, no look-ahead", False, False, 0, look=False)


if __name__ == "__main__":
    main()
