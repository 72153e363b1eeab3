"""Phase timing of every tcar_gemm_tf32 launch of one train step (debug hook tcar_debug_gemm_trace): per launch, the
mean / max over CTAs of the clock64() deltas between the 8 trace points, in microseconds at the SM clock."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.load_package()
from tcar_b200 import _native as nv, synth
from tcar_b200.model_combine import Seq2SeqAttNN

N = int(os.environ.get("N", "50000"))
content, mwdhm, _ = synth.make_catalog(N)
np.random.seed(2020)
model = Seq2SeqAttNN(dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={},
                          reverse_item=None, content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250,
                          time_hidden_size=64, l2_emb=0.0, batch_size=512, epoch=1, neg_num=20, lr=0.001, max_grad=150))
trace = torch.zeros(4096 * 8, device="cuda", dtype=torch.int64)
names = ["prologue", "tma_issue_all", "first_stage", "mma_all", "acc_ready", "epilogue", "teardown"]
MHZ = 1965.0
orig_group, orig_gemm = nv.gemm_group, nv.gemm


def report(tag, nctas_hint=None):
    torch.cuda.synchronize()
    t = trace.view(-1, 8).cpu().numpy().astype(np.float64)
    used = t[:, 0] > 0
    t = t[used]
    if len(t) == 0:
        return
    # deltas relative to the CTA start; slot 2 (TMA issue end) and 3 (first stage) belong to different warps
    rel = (t - t[:, :1]) / MHZ
    print(f"{tag:34s} ctas={len(t):4d} | " + " ".join(
        f"{n}={rel[:, i + 1].mean():5.1f}/{rel[:, i + 1].max():5.1f}" for i, n in enumerate(names)), flush=True)
    trace.zero_()


def wrap(fn, label):
    def inner(*a, **k):
        torch.cuda.synchronize()
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        torch.cuda.synchronize()
        if label == "group":
            shapes = [(q.M, q.N, sum(q.segs[i].k for i in range(q.nseg)), q.precise, q.splits) for q in a[0]]
        else:
            shapes = [(a[1], a[2], sum(s[7] for s in a[0]), k.get("precise", False), k.get("splits", 1))]
        report(f"{e0.elapsed_time(e1) * 1000:6.1f}us {shapes}"[:200])
        return r
    return inner


nv.gemm_group = wrap(orig_group, "group")
nv.gemm = wrap(orig_gemm, "gemm")
nv.lib().tcar_debug_gemm_trace(nv.ptr(trace))
for T in (2, 20):
    packed = synth.make_index_batch(N, 512, T, 20, mwdhm, seed=T)
    bt = model.to_device(torch.from_numpy(packed).pin_memory(), 512, T, 20)
    nv.lib().tcar_debug_gemm_trace(None)
    nv.gemm_group, nv.gemm = orig_group, orig_gemm
    model.train_step(bt)
    model.train_step(bt)
    torch.cuda.synchronize()
    nv.gemm_group, nv.gemm = wrap(orig_group, "group"), wrap(orig_gemm, "gemm")
    nv.lib().tcar_debug_gemm_trace(nv.ptr(trace))
    print(f"==== T = {T}  (columns: mean/max over CTAs of time since CTA start, us)")
    model.train_step(bt)
nv.lib().tcar_debug_gemm_trace(None)
