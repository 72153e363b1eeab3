"""A/B of train-step variants on ONE GPU inside one process (same clocks, same batches): GPU ms/step of the
kernel-resident loop with the look-ahead, for
  * the dense item-gradient GEMM: TMA-store epilogue (default) vs the first kernel (TCAR_BWDI_LEGACY=1),
  * the session-side backward beside that GEMM (Seq2SeqAttNN.bwd_overlap) with 148 / fewer persistent CTAs
    (TCAR_BWD_I_CTAS).
Both switches are read per call, so one model serves every variant.

    python tools/step_ab.py            (PROBE_ITEMS, PROBE_T, PROBE_STEPS as in tools/e2e_probe.py)
"""
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
    N, B, Nn = int(os.environ.get("PROBE_ITEMS", 364047)), 512, 20
    mix = [int(t) for t in os.environ.get("PROBE_T", "8,4,2,1,1,1,1,1,4,1,1,3,3,1,1,2").split(",")]
    K = int(os.environ.get("PROBE_STEPS", 48))
    content, mwdhm, _ = synth.make_catalog(N)
    np.random.seed(2020)
    args = dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={}, reverse_item=None,
                content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250, time_hidden_size=64, l2_emb=0.0,
                batch_size=B, epoch=1, neg_num=Nn, lr=0.001, max_grad=150, rank=0, world_size=1, train_parallel="dp")
    model = Seq2SeqAttNN(args)

    def batches(Ts, seed0):
        return [model.to_device(torch.from_numpy(synth.make_index_batch(N, B, t, Nn, mwdhm, seed=seed0 + i)).pin_memory(),
                                B, t, Nn) for i, t in enumerate(Ts)]

    sets = {"mix": batches(mix, 1000), "T20": batches([20] * 4, 5000)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_loop(dev):
        n = len(dev)
        best = None
        for rep in range(3):
            for i in range(4):
                model.train_step(dev[i % n], dev[(i + 1) % n])
            model.sync_updates()
            torch.cuda.synchronize()
            e0.record()
            for i in range(4, 4 + K):
                model.train_step(dev[i % n], dev[(i + 1) % n])
            model.sync_updates()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            best = ms if best is None else min(best, ms)
        return best

    variants = [("legacy bwd_i", "1", False, None), ("tma bwd_i", "0", False, None),
                ("tma + overlap, 148 CTAs", "0", True, None), ("tma + overlap, 140 CTAs", "0", True, 140),
                ("tma + overlap, 132 CTAs", "0", True, 132), ("tma + overlap, 124 CTAs", "0", True, 124),
                ("tma + overlap, 116 CTAs", "0", True, 116), ("tma, 132 CTAs, no overlap", "0", False, 132),
                ("legacy + overlap, 132 CTAs", "1", True, 132), ("tma bwd_i (again)", "0", False, None)]
    for name, legacy, overlap, ctas in variants:
        os.environ["TCAR_BWDI_LEGACY"] = legacy
        if ctas is None:
            os.environ.pop("TCAR_BWD_I_CTAS", None)
        else:
            os.environ["TCAR_BWD_I_CTAS"] = str(ctas)
        model.bwd_overlap = overlap
        out = {"variant": name}
        for key, dev in sets.items():
            out[key + "_ms"] = round(time_loop(dev), 4)
        out["loss"] = float(model.loss[:B].mean().item())
        print("step_ab " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
