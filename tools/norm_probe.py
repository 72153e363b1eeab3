import os, sys, numpy as np, torch
ROOT = "/root/repo" if os.path.exists("/root/repo/__graft_entry__.py") else os.getcwd()
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.load_package()
from tcar_b200 import synth
from tcar_b200.model_combine import Seq2SeqAttNN
for N in (364047, 20000):
    content, mwdhm, _ = synth.make_catalog(N)
    np.random.seed(2020)
    m = Seq2SeqAttNN(dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={}, reverse_item=None,
                          content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250, time_hidden_size=64, l2_emb=0.0,
                          batch_size=512, epoch=1, neg_num=20, lr=0.001, max_grad=150, rank=0, world_size=1, train_parallel="dp"))
    Ts = [8, 4, 2, 1, 1, 1, 20, 1]
    bts = [m.to_device(torch.from_numpy(synth.make_index_batch(N, 512, t, 20, mwdhm, seed=100 + i)).pin_memory(), 512, t, 20) for i, t in enumerate(Ts)]
    out = []
    for it in range(40):
        bt = bts[it % len(bts)]
        loss = m.train_step(bt)
        torch.cuda.synchronize()
        if it < 8 or it % 8 == 0:
            sm = m.ps.sqnorm_small.sum(1).sqrt().max().item()
            out.append((it, bt.T, round(float(m.ps.sqnorm_item.sqrt().item()), 3), round(sm, 3), round(float(loss.mean().item()), 3)))
    print("norm_probe N", N, "(step, T, ||g_item||, max small-tensor norm, mean loss):", out, flush=True)
    del m
    torch.cuda.empty_cache()
