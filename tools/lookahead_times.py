"""How long the two concurrent branches of the look-ahead take: table-wide Adam (side stream) and the next batch's
session forward (high-priority stream), both measured from the fork event, next to the plain serial durations."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.load_package()
from tcar_b200 import _native as nv, synth
from tcar_b200.model_combine import Seq2SeqAttNN

N = 364047
TS = [int(x) for x in os.environ.get("TS", "8,4,2,1,1,1,1,1,4,1,1,3,3,1,1,2").split(",")]
content, mwdhm, _ = synth.make_catalog(N)
np.random.seed(2020)
model = Seq2SeqAttNN(dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={},
                          reverse_item=None, content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250,
                          time_hidden_size=64, l2_emb=0.0, batch_size=512, epoch=1, neg_num=20, lr=0.001, max_grad=150))
bts = [model.to_device(torch.from_numpy(synth.make_index_batch(N, 512, T, 20, mwdhm, seed=i)).pin_memory(), 512, T, 20)
       for i, T in enumerate(TS)]
for ctas in [int(x) for x in os.environ.get("CTAS", "16,64,256").split(",")]:
    model.adam_overlap_ctas = ctas
    for i in range(4):
        model.train_step(bts[i], bts[i + 1])
    model.sync_updates(); torch.cuda.synchronize()
    model._la_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3 * len(bts)
    for i in range(n):
        model.train_step(bts[i % len(bts)], bts[(i + 1) % len(bts)])
    model.sync_updates()
    e1.record()
    torch.cuda.synchronize()
    adam = np.mean([f.elapsed_time(d) for f, d, a in model._la_events]) * 1000
    fwd = np.mean([f.elapsed_time(a) for f, d, a in model._la_events]) * 1000
    print(f"ctas/SM {ctas:4d}: step {e0.elapsed_time(e1) / n * 1000:7.1f} us | fork->adam done {adam:6.1f} us | fork->forward done {fwd:6.1f} us", flush=True)
    model._la_events = None
# serial references
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
ps, p = model.ps, nv.ptr
args = (p(ps.item), p(ps.item_m), p(ps.item_v), p(ps.item_g), p(ps.sqnorm_item), p(ps.step), 0.0, model.max_grad_f, p(ps.iext))
for c in (16, 64, 256):
    print(f"adam_item alone, {c} CTAs/SM: {timeit(lambda: nv.call('tcar_adam_item', *args, 0, ps.N + 1, None, c)):.1f} us")
k = [0]
def fwd():
    model._session_forward(bts[k[0] % len(bts)]); k[0] += 1
print(f"session forward alone (mix average): {timeit(fwd, 32):.1f} us")
