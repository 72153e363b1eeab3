"""__graft_entry__.smoke(): one tiny TCAR train step + one eval batch on cuda:0, checked against the CPU oracle."""
import numpy as np
import torch


def run(N=1500, B=64, T=5, Nn=20, verbose=True):
    from oracle import tcar_oracle as O
    from . import synth
    from .model_combine import Seq2SeqAttNN

    torch.cuda.set_device(0)
    content, mwdhm, category = synth.make_catalog(N, seed=3)
    np.random.seed(2020)
    args = dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id={i: int(category[i]) for i in range(N)},
                item_freq_dict_norm={}, reverse_item={i: i for i in range(N)}, content_emb=content,
                emb_stddev=0.002, stddev=0.05, hidden_size=250, time_hidden_size=64, l2_emb=0.0, batch_size=512,
                epoch=1, neg_num=Nn, lr=0.001, max_grad=150)
    model = Seq2SeqAttNN(args)
    params = {k: v.double() for k, v in model.ps.export().items()}
    packed = synth.make_index_batch(N, B, T, Nn, mwdhm, seed=1)
    bt = model.to_device(torch.from_numpy(packed).pin_memory(), B, T, Nn)
    batch = {k: torch.from_numpy(v) for k, v in synth.unpack(packed, B, T, Nn).items()}
    cont64, mw64 = torch.from_numpy(content).double(), torch.from_numpy(mwdhm.astype(np.int64))
    ref = O.eval_batch(params, cont64, mw64, batch, args["category_id"], args["reverse_item"])
    top, ngt, ce = model.eval_step(bt)
    torch.cuda.synchronize()
    top, ngt = top.cpu().numpy(), ngt.cpu().numpy()
    ref_rank = np.array([int((row[l] < row).sum()) + 1 for row, l in zip(ref["scores"], batch["label"].numpy())])
    agree = (top == ref["top20"]).all(1).mean()
    rank_ok = ((ngt + 1 <= 20) == (ref_rank <= 20)).all() and (np.where(ref_rank <= 20, ngt + 1 == ref_rank, True)).all()
    ce_err = np.abs(ce.cpu().numpy() - ref["cross_loss"].ravel()).max()
    loss = model.train_step(bt)
    torch.cuda.synchronize()
    out, _ = O.loss_and_grads(params, cont64, mw64, batch)
    loss_err = np.abs(loss.cpu().numpy() - out["loss"].numpy().ravel()).max()
    if verbose:
        print(f"smoke: top20 rows identical {agree:.3f}, rank parity {rank_ok}, |dCE| {ce_err:.2e}, |dloss| {loss_err:.2e}")
    assert agree >= 0.95 and rank_ok and ce_err < 3e-2 and loss_err < 3e-2, "smoke parity failed"
    return True
