"""Generate golden vectors from the REFERENCE's own source (run in the build container, where /root/reference
exists; the resulting .npz/.json fixtures are committed because the reference cannot travel to the GPU box).

  python tests/golden/make_golden.py

What is executed from /root/reference, unmodified unless stated:
  * sampler.py            -- imported directly (pure Python/NumPy): batch composition, time features, dwell buckets,
                             uniform and impression-based negatives under the reference's seeds (main.py:9-12).
  * util.py               -- imported with tests/golden/tf1_shim.py standing in for TensorFlow: cau_metrics.
  * modules.py            -- imported the same way.
  * model_combine.py      -- the file has a repeated keyword argument (`interval=None,` on line 113 and
                             `interval=seq_active_time` on line 115) and does not compile; the source is read,
                             line 113 is dropped IN MEMORY (the ablation switch) and the result is
                             exec'd.  Seq2SeqAttNN.__init__ then runs the whole graph + one Adam step eagerly.
Outputs (tests/golden/): tcar_ref_default.npz, tcar_ref_clip.npz, sampler_ref.json, metrics_ref.json
"""
import datetime
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import tf1_shim as tf  # noqa: E402

tf.install()
sys.path.insert(0, REF)

NAMES = ["item", "pos", "month", "day", "week", "hour", "minute", "dur", "W_in", "W_c", "W_i", "w_r", "Wq1", "bq1",
         "Wq2", "bq2", "W_a", "b_a", "W1", "W2", "w_t", "W_p", "b_p"]


def load_reference_model():
    import modules  # noqa: F401  (reference modules.py, through the shim)
    src = open(os.path.join(REF, "model_combine.py")).read().split("\n")
    assert src[112].strip() == "interval=None,", src[112]
    del src[112]
    mod = types.ModuleType("model_combine_ref")
    mod.__file__ = os.path.join(REF, "model_combine.py")
    exec(compile("\n".join(src), mod.__file__, "exec"), mod.__dict__)
    return mod


DENSE = NAMES[8:]          # the 15 dense tensors, each one tf.random_normal call, in creation order


def pack16(a):
    """float16 with a per-tensor scale (max |a| -> 1): 5e-4 relative per element, enough for the norm-wise 2e-2
    comparisons the big fixture is used for, at half the bytes."""
    a = np.asarray(a, dtype=np.float64)
    s = float(np.abs(a).max()) or 1.0
    return (a / s).astype(np.float16), np.float64(s)


def make_case(mod, name, N, B, T, Nn, emb_stddev, max_grad, seed, hidden=24, compact=False):
    """hidden_size is a free flag in the reference (main.py:109) as long as it equals the content width; the small
    fixtures use 24 so they stay a few hundred KB.  time_hidden_size must be 64 (modules.py:138).
    compact=True (the hidden=250 fixture the CUDA path is compared with directly): initial values are NOT stored --
    the embedding tables come from the legacy NumPy stream under seed 2020 (modules.py:32) and the dense weights from
    tf1_shim.rn_values(seed, k, ...), both reproducible from NumPy alone; checksums are stored instead -- and
    gradients / updates are stored as scaled float16."""
    rs = np.random.RandomState(seed)
    content = rs.normal(0, 0.2, (N + 1, hidden))
    big = rs.choice(np.arange(1, N + 1), N // 4, replace=False)          # a quarter of the rows get norm in (1,3]
    content[big] *= (rs.uniform(1.0, 3.0, (len(big), 1)) / np.linalg.norm(content[big], axis=1, keepdims=True))
    content[0] = 0
    content = content.astype(np.float32).astype(np.float64)
    mwdhm = np.stack([rs.randint(1, 13, N), rs.randint(1, 32, N), rs.randint(1, 8, N), rs.randint(1, 25, N),
                      rs.randint(1, 61, N)], 1).astype(np.int32)
    feed = {
        "inputs_seq": rs.randint(1, N + 1, (B, T)).astype(np.int32),
        "publish_month": rs.randint(1, 13, (B, T)).astype(np.int32),
        "publish_day": rs.randint(1, 32, (B, T)).astype(np.int32),
        "publish_week": rs.randint(1, 8, (B, T)).astype(np.int32),
        "publish_hour": rs.randint(1, 25, (B, T)).astype(np.int32),
        "publish_minute": rs.randint(1, 61, (B, T)).astype(np.int32),
        "click_week": rs.randint(0, 7, B).astype(np.int32),
        "click_hour": rs.randint(0, 24, B).astype(np.int32),
        "lab_input": rs.randint(0, N, B).astype(np.int32),
        "lab_neg": rs.randint(0, N, (B, Nn)).astype(np.int32),
        "active_time": rs.randint(0, 11, (B, T)).astype(np.int32),
        "is_training": True,
    }
    feed["inputs_seq"][0, :2] = feed["inputs_seq"][1, 0]          # duplicate ids inside the batch
    feed["lab_neg"][0, 0] = feed["lab_input"][0]                   # a negative colliding with the label
    tf.reset(seed)
    tf.FEED.update(feed)
    np.random.seed(2020)                                           # main.py:11-12, consumed by modules.embedding
    args = dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id={}, item_freq_dict_norm={}, reverse_item={},
                content_emb=content, emb_stddev=emb_stddev, stddev=0.05, hidden_size=hidden, time_hidden_size=64,
                l2_emb=0.0, batch_size=B, epoch=1, neg_num=Nn, lr=0.001, max_grad=max_grad)
    model = mod.Seq2SeqAttNN(args)
    tv = tf.STATE["trainable"]
    assert len(tv) == len(NAMES) == len(model.variables_names), (len(tv), model.variables_names)
    out = {"content": content.astype(np.float32), "mwdhm": mwdhm, "max_grad": np.float64(max_grad),
           "lr": np.float64(0.001), "var_names": np.array(model.variables_names)}
    for k, v in feed.items():
        if k != "is_training":
            out["feed_" + k] = v
    out["seed"], out["emb_stddev"], out["hidden"] = np.int64(seed), np.float64(emb_stddev), np.int64(hidden)
    for n, v, g, c in zip(NAMES, tv, tf.STATE["grads"], tf.STATE["capped"]):
        init = tf.STATE["init"][id(v)].numpy()
        assert (init.astype(np.float32).astype(np.float64) == init).all()   # variables are float32 in the reference
        delta = v.detach().numpy() - init
        if compact:
            out["initsum_" + n] = np.array([init.sum(), (init * init).sum()])
            out["grad16_" + n], out["gradscale_" + n] = pack16(g.numpy())
            if n not in DENSE or init.ndim == 1:
                out["delta16_" + n], out["deltascale_" + n] = pack16(delta)
        else:
            out["init_" + n] = init.astype(np.float32)
            out["grad_" + n] = g.numpy().astype(np.float32)                 # raw d(sum loss)/d var
            out["delta_" + n] = delta.astype(np.float32)                    # Adam update after clip_by_norm
        out["gradnorm_" + n] = np.float64(np.linalg.norm(g.numpy()))
        out["capnorm_" + n] = np.float64(np.linalg.norm(c.numpy()))         # ||clip_by_norm(g)||
        sq = tf.STATE["slices_sq"][id(v)]
        # norm tf.clip_by_norm sees: un-aggregated IndexedSlices values for lookup-only tables, else the dense norm
        out["clipnorm_" + n] = np.float64(np.sqrt(sq)) if sq is not None else out["gradnorm_" + n]
        out["sliced_" + n] = np.bool_(sq is not None)
    out["softmax_input"] = model.softmax_input.detach().numpy()
    out["cross_loss"] = model.cross_loss.detach().numpy()
    out["loss"] = model.loss.detach().numpy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "loss", out["loss"].ravel()[:3], "|g_item|", float(out["gradnorm_item"]), "clip norms TF / dense:",
          {n: (round(float(out["clipnorm_" + n]), 4), round(float(out["gradnorm_" + n]), 4)) for n in NAMES[1:8]})
    return model


def synth_sessions(rs, n_sessions, N, max_len):
    """Tiny dict-of-lists dataset in the layout util.data_partition returns."""
    len_dict, session_dict, time_dict = {}, {}, {}
    t0 = datetime.datetime(2017, 10, 1, 0, 0, 0)
    for s in range(n_sessions):
        L = int(rs.randint(1, max_len + 1))
        key = "%d_%d" % (s, L)
        session_dict[key] = [int(x) for x in rs.randint(1, N + 1, L + 1)]
        times = []
        for _ in range(L + 1):
            click = t0 + datetime.timedelta(seconds=int(rs.randint(0, 86400 * 60)))
            pub = click - datetime.timedelta(seconds=int(rs.randint(0, 86400 * 10)))
            times.append({"click_t": click, "publish_t": pub, "delta_h": 1, "active_t": int(rs.choice([0, 1, 3, 17, 600, 1023, 1500]))})
        time_dict[key] = times
        len_dict.setdefault(L, []).append(key)
    return len_dict, session_dict, time_dict


def make_sampler_golden():
    import sampler as ref_sampler
    rs = np.random.RandomState(5)
    N = 50
    len_dict, session_dict, time_dict = synth_sessions(rs, 37, N, 4)
    item_dict = {"orig%d" % i: i + 1 for i in range(N)}
    impressions = {s: ["orig%d" % int(x) for x in rs.randint(0, N + 20, 6)] for s in range(37)}
    dump = {"N": N, "len_dict": {str(k): v for k, v in len_dict.items()}, "session_dict": session_dict,
            "time_dict": {k: [{"click_t": t["click_t"].isoformat(), "publish_t": t["publish_t"].isoformat(),
                               "active_t": t["active_t"]} for t in v] for k, v in time_dict.items()},
            "impressions": {str(k): v for k, v in impressions.items()}, "bucketized": {}, "runs": []}
    for sec in [0, 1, 2, 3, 7, 8, 1022, 1023, 1024, 5000]:
        dump["bucketized"][str(sec)] = int(ref_sampler.bucketized(sec))
    for batch_size, neg_num in [(8, 3), (5, 4)]:
        random.seed(2020)
        np.random.seed(2020)
        ld = {k: list(v) for k, v in len_dict.items()}
        s = ref_sampler.Sampler(ld, session_dict, time_dict, impressions, item_dict, neg_num, batch_size=batch_size)
        batches = []
        while s.has_next():
            b_in, b_out, b_pt, b_ct, neg, gap = s.next_batch()
            batches.append({"in": b_in, "out": [int(x) for x in b_out], "pt": [list(map(list, x)) for x in b_pt],
                            "ct": [list(map(int, x)) for x in b_ct], "neg": [[int(y) for y in x] for x in neg],
                            "gap": [[int(y) for y in x] for x in gap]})
        impre = [[int(y) for y in s.neg_neighbor_from_impre(sid)] for sid in range(5)]
        dump["runs"].append({"batch_size": batch_size, "neg_num": neg_num, "batches": batches, "impre": impre,
                             "shuffled_len_dict": {str(k): v for k, v in ld.items()}})
    # eval-style sampler (no negatives): model_combine.py:261
    random.seed(2020)
    s = ref_sampler.Sampler({k: list(v) for k, v in len_dict.items()}, session_dict, time_dict, batch_size=16)
    ev = []
    while s.has_next():
        b_in, b_out, b_pt, b_ct, neg, gap = s.next_batch()
        ev.append({"in": b_in, "out": [int(x) for x in b_out], "neg": neg, "ct": [list(map(int, x)) for x in b_ct]})
    dump["eval_run"] = ev
    json.dump(dump, open(os.path.join(HERE, "sampler_ref.json"), "w"))
    print("sampler_ref.json", sum(len(r["batches"]) for r in dump["runs"]), "batches")


def make_metrics_golden(mod):
    import util as ref_util
    rs = np.random.RandomState(11)
    preds = rs.normal(size=(9, 60)).astype(np.float32)
    preds[0, 5] = preds[0, 7]                   # tie with the label
    preds[1, :30] = 1.0                         # many ties
    labels = [5, 3, 59, 0, 17, 17, 44, 2, 31]
    recall, mrr, ndcg = ref_util.cau_metrics(preds, labels, 20)
    cat = {("o%d" % i): int(rs.randint(0, 5)) for i in range(60)}
    rev = {i: "o%d" % i for i in range(60)}
    fake = types.SimpleNamespace(category_id=cat, reverse_item=rev)
    recs = [np.argsort(p).tolist()[::-1][:20] for p in preds]
    ild = [mod.Seq2SeqAttNN.getILD(fake, r) for r in recs]
    seqs = [[int(x) for x in rs.randint(1, 61, 3)] for _ in recs]
    unexp = [mod.Seq2SeqAttNN.getUnexp(fake, s, r) for s, r in zip(seqs, recs)]
    json.dump({"preds": preds.tolist(), "labels": labels, "recall": [bool(x) for x in recall],
               "mrr": [float(x) for x in mrr], "ndcg": [float(x) for x in ndcg], "category_id": cat,
               "recs": recs, "ild": ild, "seqs": seqs, "unexp": unexp},
              open(os.path.join(HERE, "metrics_ref.json"), "w"))
    print("metrics_ref.json recall", recall)


if __name__ == "__main__":
    mod = load_reference_model()
    make_case(mod, "tcar_ref_default.npz", N=300, B=6, T=4, Nn=5, emb_stddev=0.002, max_grad=150, seed=1)
    make_case(mod, "tcar_ref_clip.npz", N=257, B=7, T=3, Nn=4, emb_stddev=0.2, max_grad=0.5, seed=2)
    make_case(mod, "tcar_ref_t1.npz", N=120, B=3, T=1, Nn=2, emb_stddev=0.05, max_grad=150, seed=3)
    # the product's own width (hidden_size 250): compared DIRECTLY with the CUDA path (tests/test_gpu_golden.py)
    make_case(mod, "tcar_ref_h250.npz", N=300, B=6, T=4, Nn=5, emb_stddev=0.05, max_grad=150, seed=4, hidden=250,
              compact=True)
    make_sampler_golden()
    make_metrics_golden(mod)
