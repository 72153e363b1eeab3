"""Host utilities mirroring the reference's util.py: data_partition (pickle loader, util.py:20-57), cau_metrics
(util.py:8-18), save_model (util.py:102-106; the reference version is broken -- stray `self`, wrong key)."""
import os
import pickle
import time

import numpy as np


def cau_metrics(preds, labels, cutoff=20):
    """util.py:8-18 on a materialised score matrix (used by tests / debugging; the hot path gets the rank from
    tcar_eval_topk's n_greater instead)."""
    preds = np.asarray(preds)
    recall, mrr, ndcg = [], [], []
    for row, lab in zip(preds, labels):
        rank = int((row[lab] < row).sum()) + 1
        recall.append(rank <= cutoff)
        mrr.append(1 / rank if rank <= cutoff else 0.0)
        ndcg.append(1 / np.log2(rank + 1) if rank <= cutoff else 0.0)
    return recall, mrr, ndcg


def _load(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def data_partition(fname, foldnum, impression_path=None):
    """Same 7-tuple as util.data_partition.  `fname` = datapath + dataset + split_way.  The reference hard-codes
    /home/sansa/recsys/TCAR/data/mind/sess_impressions.mid (util.py:47); here `impression_path` (or
    <fname>sess_impressions.mid, else <fname>neighbor_<fold>.txt) supplies the neighbour dict."""
    f = str(foldnum)
    train = (_load(fname + "len_dict_train" + f + ".pkl"), _load(fname + "session_dict_train_" + f + ".pkl"),
             _load(fname + "session_time_dict_train" + f + ".pkl"))
    test = (_load(fname + "len_dict_test" + f + ".pkl"), _load(fname + "session_dict_test_" + f + ".pkl"),
            _load(fname + "session_time_dict_test" + f + ".pkl"))
    item_dict = _load(fname + "item_dict_" + f + ".txt")
    neighbor_dict = None
    for cand in (impression_path, fname + "sess_impressions.mid", fname + "neighbor_" + f + ".txt"):
        if cand and os.path.exists(cand):
            neighbor_dict = _load(cand)
            break
    if neighbor_dict is None:
        raise FileNotFoundError("no impression / neighbour file found (pass --impression_path)")
    content_emb = _load(fname + "content_weight_" + f + ".txt")
    publish_time = _load(fname + "publish_time_" + f + ".txt")
    return train, test, item_dict, neighbor_dict, content_emb, publish_time, None


def save_model(model, args, saver=None):
    """Checkpoint = parameters + Adam moments + step (torch.save)."""
    import torch
    suf = time.strftime("%Y%m%d%H%M", time.localtime()) + "-" + args["dataset"].strip("/").replace("/", "_") + \
        "-" + args["split_way"].strip("/") + "-" + str(args["foldnum"])
    os.makedirs(args["modelpath"], exist_ok=True)
    path = os.path.join(args["modelpath"], "model.ckpt-" + suf)
    model.sync_updates()
    model.sync_item_table()                  # catalog-sharded training: collect every owner's rows (and moments) first
    model.sync_optimizer_state()             # data-parallel sharded update: collect every rank's moment slices
    if getattr(model, "rank", 0) == 0:       # now every rank holds the same state; one writer
        torch.save(model.ps.state_dict(), path)
    return path


def restore_model(model, path):
    import torch
    model.sync_updates()
    # the state dict holds tensors, ints and a dict of tensors only: no pickled code is ever executed
    model.ps.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))
