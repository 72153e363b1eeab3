"""In-situ cost of every launch of the train step: a CUDA event is recorded after each C-ABI call / torch op group, and
the time between consecutive events (kernel + launch gap, warm caches, no profiler) is averaged per call site over
the session-length mix.  Serial (no look-ahead) so that the sum is the step time."""
import os, sys, collections
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.load_package()
from tcar_b200 import _native as nv, synth
from tcar_b200.model_combine import Seq2SeqAttNN

N = int(os.environ.get("N", "364047"))
TS = [int(x) for x in os.environ.get("TS", "8,4,2,1,1,1,1,1,4,1,1,3,3,1,1,2").split(",")]
content, mwdhm, _ = synth.make_catalog(N)
np.random.seed(2020)
model = Seq2SeqAttNN(dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={},
                          reverse_item=None, content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250,
                          time_hidden_size=64, l2_emb=0.0, batch_size=512, epoch=1, neg_num=20, lr=0.001, max_grad=150))
bts = [model.to_device(torch.from_numpy(synth.make_index_batch(N, 512, T, 20, mwdhm, seed=i)).pin_memory(), 512, T, 20)
       for i, T in enumerate(TS)]
for bt in bts[:4]:
    model.train_step(bt)
torch.cuda.synchronize()

events = []
seq = [0]

def mark(label):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    events.append((label, e))

def wrap_call(fn):
    def inner(name, *a):
        r = fn(name, *a)
        seq[0] += 1
        mark(f"{seq[0]:02d} {name}")
        return r
    return inner

orig_call, orig_group, orig_gemm = nv.call, nv.gemm_group, nv.gemm
nv.call = wrap_call(orig_call)
tot = collections.OrderedDict()
cnt = collections.Counter()
steps = 0
for rep in range(3):
    for bt in bts:
        events.clear()
        seq[0] = 0
        mark("start")
        model.train_step(bt)
        mark("99 end (torch tail)")
        torch.cuda.synchronize()
        for (l0, e0), (l1, e1) in zip(events[:-1], events[1:]):
            tot[l1] = tot.get(l1, 0.0) + e0.elapsed_time(e1) * 1000
            cnt[l1] += 1
        steps += 1
print(f"per-step mean over {steps} steps (session lengths {TS}); us between consecutive events")
s = 0.0
for k in sorted(tot):
    v = tot[k] / steps
    s += v
    print(f"  {k:45s} {v:8.1f}")
print(f"  {'sum':45s} {s:8.1f}")
