"""Synthetic Globo / MIND / Adressa-shaped data (there are no datasets and no network here).

  make_catalog(N)                      content [N+1,250], publish_time_MWDHM [N,5], categories
  make_index_batch(N, B, T, Nn)        one packed int32 batch straight from RNG (bench / parity tests)
  write_dataset(dir, ...)              the reference's on-disk pickle layout for main.py
"""
import datetime
import os
import pickle

import numpy as np

H = 250


def make_catalog(N, seed=2020, big_frac=0.05):
    """content ~ N(0, 0.1^2) per element with `big_frac` of the rows rescaled to norm in (1,3] so that the
    max_norm clip is exercised; row 0 is the zero pad row."""
    rs = np.random.RandomState(seed)
    content = rs.normal(0, 0.1, (N + 1, H)).astype(np.float32)
    nb = max(1, int(N * big_frac))
    big = rs.choice(np.arange(1, N + 1), nb, replace=False)
    content[big] *= (rs.uniform(1.0, 3.0, (nb, 1)) / np.linalg.norm(content[big], axis=1, keepdims=True)).astype(np.float32)
    content[0] = 0
    mwdhm = np.stack([rs.randint(1, 13, N), rs.randint(1, 32, N), rs.randint(1, 8, N), rs.randint(1, 25, N),
                      rs.randint(1, 61, N)], 1).astype(np.int32)
    category = rs.randint(0, 461, N).astype(np.int64)          # Globo has 461 categories
    return content, mwdhm, category


def zipf_items(rs, N, size, a=1.1):
    """Zipf(a) item popularity over [1, N] (inverse-CDF on a truncated power law)."""
    u = rs.random_sample(size)
    ranks = np.floor(np.exp(u * np.log(N + 1.0))).astype(np.int64)     # log-uniform ~ Zipf(1)
    if a != 1.0:
        ranks = np.floor(((N + 1.0) ** (1 - a) * u + (1 - u)) ** (1 / (1 - a))).astype(np.int64)
    return np.clip(ranks, 1, N).astype(np.int32)


def make_index_batch(N, B, T, Nn, mwdhm, seed=0):
    """Packed int32 batch in the model_combine.Batch layout: items Zipf(1.1), publish features taken from the
    clicked items' own publish times, dwell buckets in [0,10], click week/hour uniform, labels and negatives
    uniform."""
    rs = np.random.RandomState(seed)
    M = B * T
    out = np.empty(7 * M + 3 * B + B * Nn, dtype=np.int32)
    idx = out[: 7 * M].reshape(7, B, T)
    idx[0] = zipf_items(rs, N, (B, T))
    idx[1:6] = np.moveaxis(mwdhm[idx[0] - 1], -1, 0)
    idx[6] = rs.randint(0, 11, (B, T))
    out[7 * M: 7 * M + B] = rs.randint(0, 7, B)
    out[7 * M + B: 7 * M + 2 * B] = rs.randint(0, 24, B)
    out[7 * M + 2 * B: 7 * M + 3 * B] = rs.randint(0, N, B)
    if Nn:
        out[7 * M + 3 * B:] = rs.randint(0, N, B * Nn)
    return out


def unpack(packed, B, T, Nn):
    """Packed batch -> dict of int64 arrays with the oracle's key names."""
    M = B * T
    idx = packed[: 7 * M].reshape(7, B, T).astype(np.int64)
    d = {"seq": idx[0], "pm": idx[1], "pd": idx[2], "pw": idx[3], "ph": idx[4], "pmi": idx[5], "gap": idx[6],
         "cw": packed[7 * M: 7 * M + B].astype(np.int64), "ch": packed[7 * M + B: 7 * M + 2 * B].astype(np.int64),
         "label": packed[7 * M + 2 * B: 7 * M + 3 * B].astype(np.int64)}
    if Nn:
        d["neg"] = packed[7 * M + 3 * B:].reshape(B, Nn).astype(np.int64)
    return d


def make_impressions(N, n_sessions, seed=2020, mean_len=37.0, miss=0.15):
    """MIND-shaped impression lists (SURVEY 8d cfg 4): one list per session id, length ~ LogNormal with mean
    `mean_len` (at least 2), article ids uniform; a share `miss` of the entries are articles outside item_dict
    (mind_preprocess.py:275-280: the impression logs name more articles than the clicked-item dictionary), which the
    sampler's `if randomid in self.item_dict` (sampler.py:124) rejects."""
    rs = np.random.RandomState(seed + 17)
    sigma = 0.6
    mu = np.log(mean_len) - sigma * sigma / 2
    lens = np.maximum(2, rs.lognormal(mu, sigma, n_sessions).astype(np.int64))
    out = {}
    for sidx in range(n_sessions):
        ids = rs.randint(0, N, lens[sidx])
        gone = rs.random_sample(lens[sidx]) < miss
        out[sidx] = [("x%d" if g else "a%d") % int(i) for i, g in zip(ids, gone)]
    return out


def make_sessions(N, n_sessions, max_len=20, seed=2020, p_len=0.55, train=True, impressions="few"):
    """In-memory session split in the reference's dict layout: (len_dict, session_dict, session_time_dict, item_dict,
    impressions).  Lengths follow P(T) ~ p_len^T on [1, max_len], items Zipf(1.1).
    impressions="few": 8-entry lists for the first 64 sessions (the Globo runs only need a non-empty neighbour dict);
    "mind": make_impressions() for every session."""
    rs = np.random.RandomState(seed)
    t0 = datetime.datetime(2017, 10, 1)
    publish_dt = [t0 + datetime.timedelta(minutes=int(m)) for m in rs.randint(0, 60 * 24 * 30, N)]
    pr = p_len ** np.arange(1, max_len + 1)
    lens = rs.choice(np.arange(1, max_len + 1), size=n_sessions, p=pr / pr.sum())
    len_dict, sdict, tdict = {}, {}, {}
    for sidx in range(n_sessions):
        L = int(lens[sidx])
        items = zipf_items(rs, N, L + 1).tolist()
        key = "%d_%d" % (sidx, L) if train else sidx
        sdict[key] = items
        start = t0 + datetime.timedelta(days=31, seconds=int(rs.randint(0, 86400 * 14)))
        tdict[key] = [{"click_t": start + datetime.timedelta(seconds=60 * j), "publish_t": publish_dt[it - 1],
                       "delta_h": 1, "active_t": int(np.exp(rs.uniform(0, np.log(1023))))}
                      for j, it in enumerate(items)]
        len_dict.setdefault(L, []).append(key)
    item_dict = {"a%d" % i: i + 1 for i in range(N)}
    if impressions == "mind":
        impr = make_impressions(N, n_sessions, seed)
    else:
        impr = {sidx: ["a%d" % int(x) for x in rs.randint(0, N, 8)] for sidx in range(min(n_sessions, 64))}
    return len_dict, sdict, tdict, item_dict, impr


def write_dataset(root, N=2000, n_train=5000, n_test=600, max_len=8, fold=0, seed=2020):
    """Write <root>/{len_dict,session_dict,session_time_dict}_{train,test}*.pkl, item_dict, content_weight,
    publish_time, item_freq_dict_norm, train/test_session, sess_impressions.mid and articles_category.pkl in the
    layout util.data_partition / main.load_datas read."""
    os.makedirs(root, exist_ok=True)
    rs = np.random.RandomState(seed)
    content, mwdhm, category = make_catalog(N, seed)
    t0 = datetime.datetime(2017, 10, 1)
    publish_dt = [t0 + datetime.timedelta(minutes=int(m)) for m in rs.randint(0, 60 * 24 * 30, N)]
    mw = np.array([[d.month, d.day, d.isoweekday(), d.hour + 1, d.minute + 1] for d in publish_dt], dtype=np.int32)
    item_dict = {"a%d" % i: i + 1 for i in range(N)}
    f = str(fold)

    def sessions(n, tag, train):
        len_dict, sdict, tdict, raw = {}, {}, {}, ([], [], [], [])
        for s in range(n):
            L = int(min(max_len, 1 + rs.geometric(0.45)))
            items = zipf_items(rs, N, L + 1).tolist()
            key = "%d_%d" % (s, L) if train else s
            sdict[key] = items
            start = t0 + datetime.timedelta(days=31, seconds=int(rs.randint(0, 86400 * 14)))
            ts = []
            for j, it in enumerate(items):
                click = start + datetime.timedelta(seconds=60 * j + int(rs.randint(0, 50)))
                ts.append({"click_t": click, "publish_t": publish_dt[it - 1], "delta_h": 1,
                           "active_t": int(np.exp(rs.uniform(0, np.log(1023))))})
            tdict[key] = ts
            len_dict.setdefault(L, []).append(key)
            raw[0].append(s); raw[1].append(items[:-1]); raw[2].append([0] * L); raw[3].append(items[-1])
        pickle.dump(len_dict, open(os.path.join(root, "len_dict_%s%s.pkl" % (tag, f)), "wb"))
        pickle.dump(sdict, open(os.path.join(root, "session_dict_%s_%s.pkl" % (tag, f)), "wb"))
        pickle.dump(tdict, open(os.path.join(root, "session_time_dict_%s%s.pkl" % (tag, f)), "wb"))
        pickle.dump(raw, open(os.path.join(root, "%s_session_%s.txt" % (tag, f)), "wb"))
        return n

    sessions(n_train, "train", True)
    sessions(n_test, "test", False)
    pickle.dump(item_dict, open(os.path.join(root, "item_dict_%s.txt" % f), "wb"))
    pickle.dump(content, open(os.path.join(root, "content_weight_%s.txt" % f), "wb"))
    pickle.dump((publish_dt, mw), open(os.path.join(root, "publish_time_%s.txt" % f), "wb"))
    pickle.dump({i: 1.0 / N for i in range(N)}, open(os.path.join(root, "item_freq_dict_norm_%s.txt" % f), "wb"))
    impressions = {s: ["a%d" % int(x) for x in rs.randint(0, N, 30)] for s in range(n_train)}
    pickle.dump(impressions, open(os.path.join(root, "sess_impressions.mid"), "wb"))
    pickle.dump({"a%d" % i: int(category[i]) for i in range(N)}, open(os.path.join(root, "articles_category.pkl"), "wb"))
    return root
