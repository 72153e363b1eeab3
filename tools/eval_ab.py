"""A/B of the evaluation step on one GPU: softmax guard on/off x certified top-20 on/off, and per-phase CUDA-event
times of one eval step.    python tools/eval_ab.py [--items N]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from tcar_b200 import _native as nv, synth  # noqa: E402
from tcar_b200.model_combine import Seq2SeqAttNN  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=364047)
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    N, B = a.items, 512
    content, mwdhm, _ = synth.make_catalog(N)
    np.random.seed(2020)
    model = Seq2SeqAttNN(dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={},
                              reverse_item=None, content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250,
                              time_hidden_size=64, l2_emb=0.0, batch_size=B, epoch=1, neg_num=20, lr=0.001, max_grad=150))
    rs = np.random.RandomState(2020)
    pr = 0.55 ** np.arange(1, 21)
    Ts = [int(t) for t in rs.choice(np.arange(1, 21), size=16, p=pr / pr.sum())]
    bts = [model.to_device(torch.from_numpy(synth.make_index_batch(N, B, t, 0, mwdhm, seed=77 + i)).pin_memory(), B, t, 0)
           for i, t in enumerate(Ts)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def loop(look):
        for i in range(5):
            model.eval_step(bts[i % 16], next_bt=bts[(i + 1) % 16] if look else None)
        model.sync_updates()
        torch.cuda.synchronize()
        e0.record()
        for i in range(5, 5 + a.steps):
            model.eval_step(bts[i % 16], next_bt=bts[(i + 1) % 16] if look else None)
        model.sync_updates()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps * 1e3

    for guard in (False, True):
        for cert in (False, True):
            for look in (False, True):
                model.softmax_guard, model.eval_certify = guard, cert
                print(f"guard={int(guard)} certify={int(cert)} lookahead={int(look)}: {loop(look):7.1f} us/step", flush=True)
    model.softmax_guard = model.eval_certify = True
    for ws_ in (False, True):
        model.eval_warp_select = ws_
        print(f"guard=1 certify=1 lookahead=1 warp_select={int(ws_)}: {loop(True):7.1f} us/step", flush=True)
    # phases of one step (no look-ahead), events between the calls
    model.softmax_guard = model.eval_certify = True
    shares = []
    for bt in bts:
        model.eval_step(bt)
        shares.append(int(model.uncertain[:B].sum().item()))
    print("queries sent to the widening pass per batch of 512 (T per batch", Ts, "):", shares, flush=True)
    eps = (model.top_scores[:B, 19] - model.tau[:B]).cpu().numpy()
    s = model.top_scores[:B].cpu().numpy()
    print("last batch: eps median %.3e, s20 median %.3e, s1-s20 median %.3e, s19-s20 median %.3e, cat_stats %s"
          % (np.median(eps), np.median(s[:, 19]), np.median(s[:, 0] - s[:, 19]), np.median(s[:, 18] - s[:, 19]),
             model.cat_stats.cpu().numpy()), flush=True)
    orig = nv.counted_call
    marks = []

    def traced(name, n, *args):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, ev))
        return orig(name, n, *args)

    nv.counted_call = traced
    import tcar_b200.model_combine as mc
    mc.nv.counted_call = traced
    for rep in range(3):
        marks.clear()
        model.eval_step(bts[3])
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append(("end", ev))
        torch.cuda.synchronize()
    for (n0, a0), (n1, a1) in zip(marks, marks[1:]):
        print(f"  {n0:28s} {a0.elapsed_time(a1) * 1e3:7.1f} us")


if __name__ == "__main__":
    main()
