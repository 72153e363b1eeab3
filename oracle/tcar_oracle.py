"""CPU ORACLE (test infrastructure only -- never imported by the product path).

A restatement, in plain torch-on-CPU (fp32 or fp64, autograd for the gradients), of the TensorFlow-1 graph that
summmeer/session-based-news-recommendation builds for TCAR, plus its optimiser and evaluation metrics:

    model_combine.py:52-147   graph (embeddings, two attention poolings, scoring matmul, losses)
    modules.py:13-152         embedding / linear_2d / linear_3d / count_alpha_* / *_attention_layer
    util.py:92-100            normalizer (un-stabilised softmax + 1e-9)
    model_combine.py:151-163  tf.train.AdamOptimizer + per-tensor tf.clip_by_norm
    util.py:8-18              cau_metrics;  model_combine.py:174-194 getILD / getUnexp; :283-314 eval loop

Pinning status: the reference's arithmetic lives in TensorFlow 1.x (un-vendored, un-pinned, not installable here)
and the reference has no tests or golden vectors.  This oracle is pinned against the reference's OWN graph-building
code (model_combine.py / modules.py executed unmodified except for the repeated-keyword line 113, through the
eager TF1 shim in tests/golden/tf1_shim.py) by tests/golden/make_golden.py -> tests/golden/tcar_ref_*.npz, and
against the reference's own sampler.py / util.cau_metrics / getILD / getUnexp run directly.  What stays
"parity unpinned" is the TensorFlow kernel boundary itself (embedding_lookup(max_norm), sparse softmax CE,
clip_by_norm, Adam), which is restated from the TF 1.x documented semantics in the shim and here.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

H, TH = 250, 64
# creation order of tf.trainable_variables() in model_combine.py:52-127 (the order Adam / clip iterate over)
PARAM_ORDER = [
    "item", "pos", "month", "day", "week", "hour", "minute", "dur",
    "W_in", "W_c", "W_i", "w_r", "Wq1", "bq1", "Wq2", "bq2", "W_a", "b_a",
    "W1", "W2", "w_t", "W_p", "b_p",
]
TIME_ROWS = {"month": 13, "day": 32, "week": 8, "hour": 25, "minute": 61}


def param_shapes(n_items: int) -> Dict[str, Tuple[int, ...]]:
    """Shapes of the 23 trainable tensors; `n_items` = N (table has N+1 rows)."""
    return {
        "item": (n_items + 1, H), "pos": (40, H),
        "month": (13, TH), "day": (32, TH), "week": (8, TH), "hour": (25, TH), "minute": (61, TH), "dur": (11, TH),
        "W_in": (2 * H, H), "W_c": (H, H), "W_i": (TH, H), "w_r": (H, 1),
        "Wq1": (2 * TH, H), "bq1": (H,), "Wq2": (H, 2 * H), "bq2": (2 * H,),
        "W_a": (2 * H, 2 * H), "b_a": (2 * H,),
        "W1": (5 * TH, H), "W2": (H, H), "w_t": (H, 1),
        "W_p": (5 * TH, 5 * TH), "b_p": (5 * TH,),
    }


def init_params(n_items: int, emb_stddev: float = 0.002, stddev: float = 0.05, seed: int = 2020,
                dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reference initialisers.  Embedding tables: legacy np.random.seed(seed) then np.random.normal draws in
    creation order item -> pos -> month -> day -> week -> hour -> minute -> duration (modules.py:32,
    model_combine.py:54-107; pos uses the default stddev 0.02 and no zero pad, duration no zero pad).
    Dense weights: N(0, stddev) (modules.py:50-51,65) -- TF's tf.random_normal stream is not reproducible
    outside TF, so these come from a numpy Generator seeded with `seed`."""
    rs = np.random.RandomState(seed)
    shapes = param_shapes(n_items)
    p = {}
    for name, sd, zero_pad in [("item", emb_stddev, True), ("pos", 0.02, False), ("month", emb_stddev, True),
                               ("day", emb_stddev, True), ("week", emb_stddev, True), ("hour", emb_stddev, True),
                               ("minute", emb_stddev, True), ("dur", emb_stddev, False)]:
        t = rs.normal(0, sd, shapes[name])
        if zero_pad:
            t[0] = 0.0
        p[name] = torch.tensor(t.astype(np.float32), dtype=dtype)
    g = np.random.default_rng(seed)
    for name in PARAM_ORDER[8:]:
        p[name] = torch.tensor(g.normal(0, stddev, shapes[name]).astype(np.float32), dtype=dtype)
    return p


# ----------------------------------------------------------------------------------------------- building blocks
def clip_rows(x: torch.Tensor) -> torch.Tensor:
    """tf.nn.embedding_lookup(..., max_norm=1) == clip_by_norm over the last axis (modules.py:36):
    y = x * 1 / max(||x||, 1); zero rows pass through; differentiable through the norm."""
    sq = (x * x).sum(-1, keepdim=True)
    safe = torch.where(sq > 0, sq, torch.ones_like(sq))
    norm = torch.where(sq > 0, safe.sqrt(), sq)
    return x / torch.clamp(norm, min=1.0)


def embedding_lookup(table: torch.Tensor, ids: torch.Tensor, max_norm: bool = True, tap=None) -> torch.Tensor:
    """tap (optional list): collects the gathered rows, whose gradients are the VALUES of the tf.IndexedSlices
    gradient this lookup contributes (one slice per looked-up id, duplicates not summed)."""
    rows = table[ids.long()]
    if tap is not None:
        tap.append(rows)
    return clip_rows(rows) if max_norm else rows


def normalizer(x: torch.Tensor, axis: int = 1) -> torch.Tensor:
    """util.py:92-100: exp(x) / (sum exp(x) + 1e-9), no max subtraction."""
    e = torch.exp(x)
    return e / (e.sum(axis, keepdim=True) + 1e-9)


def linear_3d(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """modules.py:57-70 with active=None (every TCAR call site): x @ w, no bias."""
    return x @ w


def linear_2d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, active: str = "tanh") -> torch.Tensor:
    """modules.py:43-55: act(x @ w + b)."""
    r = x @ w + b
    return {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[active](r)


def count_alpha_m(p, X, C, D, ct):
    """modules.py:120-152 (active='sigmoid' default): nrm(sigmoid(XW+CW+DW) w) + nrm(X . q)."""
    res = linear_3d(X, p["W_in"]) + linear_3d(C, p["W_c"]) + linear_3d(D, p["W_i"])
    e1 = linear_3d(torch.sigmoid(res), p["w_r"]).squeeze(-1)
    alpha = normalizer(e1)
    q = linear_2d(ct, p["Wq1"], p["bq1"], "relu")
    q = linear_2d(q, p["Wq2"], p["bq2"], "tanh")
    e2 = torch.matmul(X, q.unsqueeze(-1))                      # [B,T,1]
    alpha2 = normalizer(e2).squeeze(-1)
    return alpha + alpha2


def count_alpha_s(p, P, C):
    """modules.py:86-101."""
    res = linear_3d(P, p["W1"]) + linear_3d(C, p["W2"])
    e = linear_3d(torch.sigmoid(res), p["w_t"]).squeeze(-1)
    return normalizer(e)


LOOKUP_ONLY = ("pos", "month", "day", "week", "hour", "minute", "dur")   # read through embedding_lookup only


def forward(p: Dict[str, torch.Tensor], content: torch.Tensor, mwdhm: torch.Tensor, batch: Dict[str, torch.Tensor],
            want_scores: bool = True, taps: Optional[Dict[str, list]] = None) -> Dict[str, torch.Tensor]:
    """The TCAR graph, model_combine.py:52-147, in evaluation order.

    batch: seq [B,T] (1-based), pm pd pw ph pmi [B,T], cw ch [B], gap [B,T], label [B] (0-based),
           neg [B,Nn] (0-based, optional).  mwdhm [N,5] int.  content [N+1,250] frozen."""
    seq = batch["seq"].long()
    B, T = seq.shape

    def look(name, ids):
        return embedding_lookup(p[name], ids, tap=None if taps is None else taps.setdefault(name, []))

    # the reference looks dec_pos up with a tiled [B,T] index (model_combine.py:57): B*T slices, not T
    pos_ids = torch.arange(T).unsqueeze(0).expand(B, T)
    E_i = embedding_lookup(p["item"], seq) + look("pos", pos_ids)                                       # :54-65
    E_c = embedding_lookup(content, seq)                                                                # :67-68
    P = torch.cat([look("month", batch["pm"]), look("day", batch["pd"]), look("week", batch["pw"]),
                   look("hour", batch["ph"]), look("minute", batch["pmi"])], -1)                        # :73-84
    cand_t = torch.cat([look("month", mwdhm[:, 0]), look("day", mwdhm[:, 1]), look("week", mwdhm[:, 2]),
                        look("hour", mwdhm[:, 3]), look("minute", mwdhm[:, 4])], -1)                    # :86-92
    ct = torch.cat([look("week", batch["cw"]), look("hour", batch["ch"])], -1)                          # :94-97
    D = look("dur", batch["gap"])                                                                       # :106-107
    X = torch.cat([E_i, E_c], -1)                                                                       # :111
    alpha = count_alpha_m(p, X, E_c, D, ct)                                                             # :112-117
    pooled = torch.matmul(alpha.unsqueeze(1), X).squeeze(1)                                             # modules.py:116-117
    a_ic = linear_2d(pooled, p["W_a"], p["b_a"])                                                        # :119
    alpha_t = count_alpha_s(p, P, E_c)                                                                  # :124-125
    pooled_t = torch.matmul(alpha_t.unsqueeze(1), P).squeeze(1)
    a_pt = linear_2d(pooled_t, p["W_p"], p["b_p"])                                                      # :127
    attout = torch.cat([a_ic, a_pt], -1)                                                                # :132
    items_ic = torch.cat([p["item"][1:], content[1:]], -1)                                              # :135 (unclipped)
    items = torch.cat([items_ic, cand_t], -1)                                                           # :136
    S = attout @ items.t()                                                                              # :138
    out = {"a_ic": a_ic, "a_pt": a_pt, "alpha": alpha, "alpha_t": alpha_t, "X": X, "P": P, "D": D, "ct": ct,
           "pooled": pooled, "pooled_t": pooled_t}
    label = batch["label"].long()
    lse = torch.logsumexp(S, dim=1)
    out["cross_loss"] = (lse - S.gather(1, label[:, None]).squeeze(1)).unsqueeze(1)                     # :145
    if want_scores:
        out["softmax_input"] = S
    if "neg" in batch and batch["neg"] is not None:
        neg = batch["neg"].long()
        neg_logits = torch.matmul(items_ic[neg], a_ic.unsqueeze(-1)).sum(1)                             # :142 [B,1]
        neg_fb = -torch.log(1 - torch.sigmoid(neg_logits) + 1e-24)                                      # :143
        out["neg_feedback"] = neg_fb
        out["loss"] = out["cross_loss"] + 0.01 * neg_fb                                                 # :147
    return out


# ----------------------------------------------------------------------------------------------- optimiser
def clip_by_norm(g: torch.Tensor, clip: float) -> torch.Tensor:
    """tf.clip_by_norm: g * clip / max(||g||, clip)  (model_combine.py:158-160)."""
    n = torch.sqrt((g * g).sum())
    return g * clip / torch.clamp(n, min=clip)


class TFAdam:
    """tf.train.AdamOptimizer(lr) with TF defaults b1=.9 b2=.999 eps=1e-8 and the TF update rule
    lr_t = lr sqrt(1-b2^t)/(1-b1^t);  m += (g-m)(1-b1);  v += (g^2-v)(1-b2);  p -= lr_t m / (sqrt(v)+eps)."""

    def __init__(self, params: Dict[str, torch.Tensor], lr: float):
        self.lr, self.t = lr, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor]) -> None:
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - 0.999 ** self.t) / (1 - 0.9 ** self.t)
        for k in PARAM_ORDER:
            g = grads[k]
            self.m[k] += (g - self.m[k]) * (1 - 0.9)
            self.v[k] += (g * g - self.v[k]) * (1 - 0.999)
            params[k] -= lr_t * self.m[k] / (self.v[k].sqrt() + 1e-8)


def loss_and_grads(p, content, mwdhm, batch, tf_slice_norms: bool = False):
    """Gradients of sum_b loss_b (optimizer.compute_gradients on a [B,1] tensor sums it; model_combine.py:156).

    tf_slice_norms=True additionally returns, per tensor, the norm tf.clip_by_norm (model_combine.py:158-160) sees
    in TensorFlow 1.x: for the seven tables read only through embedding_lookup (LOOKUP_ONLY) the gradient is a
    tf.IndexedSlices whose values are the per-lookup slices, CONCATENATED over the lookups (publish time of the
    clicks, of the N candidates, click context) and not summed per row, and clip_by_norm norms those values; every
    other tensor (item_emb included: it also feeds the dense [1:] slice) has a dense gradient."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    taps: Dict[str, list] = {} if tf_slice_norms else None
    out = forward(leaves, content, mwdhm, batch, want_scores=False, taps=taps)
    if tf_slice_norms:
        for rows in (r for lst in taps.values() for r in lst):
            rows.retain_grad()
    out["loss"].sum().backward()
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])) for k in PARAM_ORDER}
    outs = {k: v.detach() for k, v in out.items()}
    if not tf_slice_norms:
        return outs, grads
    norms = {k: float(torch.sqrt((g.double() ** 2).sum())) for k, g in grads.items()}
    for name in LOOKUP_ONLY:
        norms[name] = math.sqrt(sum(float((r.grad.double() ** 2).sum()) for r in taps[name] if r.grad is not None))
    return outs, grads, norms


def train_step(p, adam: TFAdam, content, mwdhm, batch, max_grad: Optional[float] = 150.0,
               tf_slice_norms: bool = False):
    """One sess.run([loss, global_step, train_op]) (model_combine.py:231-234); updates `p` in place.

    Clip norm: by default the norm of the aggregated (dense) gradient of every tensor -- what the CUDA path computes.
    tf_slice_norms=True uses TensorFlow's reading for the lookup-only tables (see loss_and_grads).  The two agree
    whenever neither norm exceeds max_grad (then nothing is clipped); when a clip fires on one of those seven tables
    they differ in the clip FACTOR only.  This is a documented deviation of the product ("Oracle")."""
    if tf_slice_norms:
        out, grads, norms = loss_and_grads(p, content, mwdhm, batch, tf_slice_norms=True)
    else:
        out, grads = loss_and_grads(p, content, mwdhm, batch)
    if max_grad is not None:
        if tf_slice_norms:
            grads = {k: g * float(max_grad) / max(norms[k], float(max_grad)) for k, g in grads.items()}
        else:
            grads = {k: clip_by_norm(g, float(max_grad)) for k, g in grads.items()}
    adam.step(p, grads)
    return out, grads


# ----------------------------------------------------------------------------------------------- evaluation
def cau_metrics(preds: np.ndarray, labels: Sequence[int], cutoff: int = 20):
    """util.py:8-18: rank = #(S > S[label]) + 1 (strict: the label wins ties)."""
    recall, mrr, ndcg = [], [], []
    for row, lab in zip(preds, labels):
        rank = int((row[lab] < row).sum()) + 1
        recall.append(rank <= cutoff)
        mrr.append(1 / rank if rank <= cutoff else 0.0)
        ndcg.append(1 / np.log2(rank + 1) if rank <= cutoff else 0.0)
    return recall, mrr, ndcg


def top20(preds: np.ndarray, k: int = 20) -> np.ndarray:
    """model_combine.py:301 `np.argsort(pred)[::-1][:20]`, with the tie order DEFINED as lower item id first
    (the reference's is implementation-defined; identical on tie-free rows)."""
    ids = np.arange(preds.shape[1])
    return np.stack([np.lexsort((ids, -row))[:k] for row in preds])


def get_ild(rec: Sequence[int], category_id, reverse_item) -> float:
    """model_combine.py:174-182."""
    n = len(rec)
    score = 0
    for i in range(n):
        for j in range(n):
            if j != i and category_id[reverse_item[rec[i]]] != category_id[reverse_item[rec[j]]]:
                score += 1
    return score / (n * (n - 1))


def get_unexp(in_seq: Sequence[int], rec: Sequence[int], category_id, reverse_item) -> float:
    """model_combine.py:184-194."""
    n = len(rec)
    if n == 0:
        return 0
    score = 0
    for i in range(n):
        for ini in in_seq:
            if category_id[reverse_item[rec[i]]] != category_id[reverse_item[ini - 1]]:
                score += 1
    return score / (n * len(in_seq))


def eval_batch(p, content, mwdhm, batch, category_id, reverse_item):
    """One iteration of the loop at model_combine.py:264-306 for a single batch."""
    with torch.no_grad():
        out = forward(p, content, mwdhm, batch)
    S = out["softmax_input"].numpy()
    labels = batch["label"].numpy()
    recall, mrr, ndcg = cau_metrics(S, labels, 20)
    tops = top20(S)
    ild = [get_ild(list(t), category_id, reverse_item) for t in tops]
    unexp = [get_unexp(list(batch["seq"][i].numpy()), list(t), category_id, reverse_item) for i, t in enumerate(tops)]
    return {"recall": recall, "mrr": mrr, "ndcg": ndcg, "top20": tops, "ild": ild, "unexp": unexp,
            "cross_loss": out["cross_loss"].numpy(), "scores": S}


# ----------------------------------------------------------------------------------------------- sharded evaluation
def merge_topk(ids: np.ndarray, scores: np.ndarray, k: int = 20):
    """Merge G per-shard top-k lists [G,B,k] into the global top-k ordered by (score desc, id asc); id < 0 marks an
    empty slot.  This is the defined order of top20() above applied to the union of the shard lists, i.e. what
    `np.argsort(pred)[::-1][:20]` (model_combine.py:301) yields on tie-free rows of the unsharded score matrix."""
    G, B, kk = ids.shape
    out_i = np.full((B, k), -1, dtype=np.int32)
    out_s = np.full((B, k), -np.inf, dtype=np.float32)
    for b in range(B):
        i = ids[:, b].reshape(-1).astype(np.int64)
        s = scores[:, b].reshape(-1).astype(np.float32)
        keep = i >= 0
        i, s = i[keep], s[keep]
        order = np.lexsort((i, -s))[:k]
        out_i[b, : len(order)] = i[order]
        out_s[b, : len(order)] = s[order]
    return out_i, out_s


EXP_LIMIT2 = 80.0          # TCAR_EXP_LIMIT2 (include/tcar_b200.h)


def merge_eval_blocks(blocks: np.ndarray, B: int, k: int = 20):
    """Reference of tcar_eval_merge: G result blocks (layout TCAR_EVAL_OFF_*: scores [512,k] f32 | ids [512,k] i32 |
    n_greater [512] i32 | sumexp [512] f32 | rowmax [512] f32, as float32 words) -> global top-k (score desc, id asc),
    summed rank counts (util.py:14) and CE = logsumexp(S) - S[label] (model_combine.py:145) from the shards' partial
    sums, each relative to its own exponent shift (rowmax, log2 units, applied only above EXP_LIMIT2)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.float32)
    G = blocks.shape[0]
    bi = blocks.view(np.int32)
    Q = 512
    sc = blocks[:, : Q * k].reshape(G, Q, k)[:, :B]
    ids = bi[:, Q * k: 2 * Q * k].reshape(G, Q, k)[:, :B]
    ngt = bi[:, 2 * Q * k: 2 * Q * k + Q][:, :B].astype(np.int64).sum(0)
    sumexp = blocks[:, 2 * Q * k + Q: 2 * Q * k + 2 * Q][:, :B].astype(np.float64)
    rowmax = blocks[:, 2 * Q * k + 2 * Q: 2 * Q * k + 3 * Q][:, :B].astype(np.float64)
    shift = np.where(rowmax > EXP_LIMIT2, rowmax, 0.0)
    M = shift.max(0)
    ce = np.log((sumexp * np.exp2(shift - M)).sum(0)) + M * math.log(2.0)
    out_i, out_s = merge_topk(ids, sc, k)
    return out_i, out_s, ngt, ce
