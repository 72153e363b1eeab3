"""Timeline of the single-GPU evaluation step with the one-batch look-ahead: when (relative to the step's first launch)
the scoring GEMM, the softmax sums + guard pass, the top-20 selection and the NEXT batch's session forward (high-priority
stream) finish.  Averages over the session-length mix.

    python tools/eval_probe.py
"""
import collections
import json
import os
import sys

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
    N, B = int(os.environ.get("PROBE_ITEMS", 364047)), 512
    Ts = [int(t) for t in os.environ.get("PROBE_T", "8,4,2,1,1,1,1,1,4,1,1,3,3,1,1,2").split(",")]
    content, mwdhm, _ = synth.make_catalog(N)
    np.random.seed(2020)
    model = Seq2SeqAttNN(dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={},
                              reverse_item=None, content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250,
                              time_hidden_size=64, l2_emb=0.0, batch_size=B, epoch=1, neg_num=20, lr=0.001, max_grad=150,
                              rank=0, world_size=1, train_parallel="dp"))
    dev = [model.to_device(torch.from_numpy(synth.make_index_batch(N, B, t, 0, mwdhm, seed=77 + i)).pin_memory(), B, t, 0)
           for i, t in enumerate(Ts)]
    n = len(dev)
    for look in (True, False):
        run(model, dev, look)


def run(model, dev, look):
    n = len(dev)
    nxt = (lambda i: dev[(i + 1) % n]) if look else (lambda i: None)
    for i in range(8):
        model.eval_step(dev[i % n], next_bt=nxt(i))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 64
    e0.record()
    for i in range(K):
        model.eval_step(dev[i % n], next_bt=nxt(i))
    model.sync_updates()
    e1.record()
    torch.cuda.synchronize()
    print("eval_probe " + json.dumps({"look_ahead": look, "ms_per_step": round(e0.elapsed_time(e1) / K, 4)}), flush=True)
    tot, cnt = collections.OrderedDict(), 0
    for i in range(K):
        model._eval_trace = []
        model.eval_step(dev[i % n], next_bt=nxt(i))
        torch.cuda.synchronize()
        tr = model._eval_trace
        t0 = tr[0][1]
        for name, ev in tr[1:]:
            tot[name] = tot.get(name, 0.0) + t0.elapsed_time(ev) * 1e3
        cnt += 1
    model._eval_trace = None
    print("eval_probe (us after the step's first launch, each step synchronised) " +
          json.dumps({k: round(v / cnt, 1) for k, v in tot.items()}), flush=True)


if __name__ == "__main__":
    main()
