"""The CUDA path against numbers produced by RUNNING THE REFERENCE (tests/golden/tcar_ref_h250.npz: the reference's own
model_combine.py / modules.py executed by tests/golden/make_golden.py at hidden_size = 250, the width the kernels are
built for) -- no oracle in between.  Initial values are regenerated from NumPy streams (the fixture stores checksums):
embedding tables from the legacy global stream under seed 2020 in creation order (modules.py:32), dense weights from
tf1_shim.rn_values(seed, k, ...).

Tolerances: scores / losses through the bf16 scoring GEMM rtol 5e-3 / atol 2e-2; gradients norm-wise 2e-2 (the fixture
stores them as scaled float16, 5e-4 per element); the Adam update of step 1 is -lr * g / (|g| + eps sqrt(1-b2)) ~
-lr * sign(g): elements whose reference gradient is not tiny must move by the same amount (2 % of lr)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, G)

NAMES = ["item", "pos", "month", "day", "week", "hour", "minute", "dur", "W_in", "W_c", "W_i", "w_r", "Wq1", "bq1",
         "Wq2", "bq2", "W_a", "b_a", "W1", "W2", "w_t", "W_p", "b_p"]


def load_fixture():
    import tf1_shim
    from tcar_b200.params import SMALL
    z = np.load(os.path.join(G, "tcar_ref_h250.npz"))
    N, seed, emb_sd = z["content"].shape[0] - 1, int(z["seed"]), float(z["emb_stddev"])
    shapes = dict(SMALL)
    shapes["item"] = (N + 1, 250)
    init = {}
    np.random.seed(2020)                                                    # main.py:11-12
    for name, sd, zero_pad in [("item", emb_sd, True), ("pos", 0.02, False), ("month", emb_sd, True),
                               ("day", emb_sd, True), ("week", emb_sd, True), ("hour", emb_sd, True),
                               ("minute", emb_sd, True), ("dur", emb_sd, False)]:
        t = np.random.normal(0, sd, shapes[name])                           # modules.py:32-35
        if zero_pad:
            t[0] = 0.0
        init[name] = t.astype(np.float32)
    for k, name in enumerate(NAMES[8:]):
        init[name] = tf1_shim.rn_values(seed, k, shapes[name], 0.05).astype(np.float32).reshape(shapes[name])
    for name in NAMES:
        a = init[name].astype(np.float64)
        np.testing.assert_allclose([a.sum(), (a * a).sum()], z["initsum_" + name], rtol=1e-12, atol=1e-12,
                                   err_msg=f"regenerated initial value of {name} differs from the reference run")
    return z, init, N


def grad_of(z, name):
    return z["grad16_" + name].astype(np.float64) * float(z["gradscale_" + name])


def build_model(z, init, N):
    from tcar_b200.model_combine import Seq2SeqAttNN
    from tcar_b200.sampler import pack_batch
    B, T = z["feed_inputs_seq"].shape
    Nn = z["feed_lab_neg"].shape[1]
    args = dict(publish_time_MWDHM=z["mwdhm"], itemnum=N, category_id=None, item_freq_dict_norm={}, reverse_item=None,
                content_emb=z["content"], emb_stddev=float(z["emb_stddev"]), stddev=0.05, hidden_size=250,
                time_hidden_size=64, l2_emb=0.0, batch_size=512, epoch=1, neg_num=Nn, lr=float(z["lr"]),
                max_grad=float(z["max_grad"]))
    model = Seq2SeqAttNN(args)
    model.ps.load(init)
    pt = [z["feed_publish_" + k].tolist() for k in ("month", "day", "week", "hour", "minute")]
    ct = [None, None, z["feed_click_week"].tolist(), z["feed_click_hour"].tolist(), None]
    packed, B, T, Nn = pack_batch(z["feed_inputs_seq"].tolist(), z["feed_lab_input"].tolist(), pt, ct,
                                  z["feed_lab_neg"].tolist(), z["feed_active_time"].tolist())
    bt = model.to_device(torch.from_numpy(packed).pin_memory(), B, T, Nn)
    return model, bt, B


def test_forward_losses_and_scores_match_the_reference_run():
    z, init, N = load_fixture()
    model, bt, B = build_model(z, init, N)
    loss, ce = model.forward_train(bt)
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss.cpu().numpy(), z["loss"].ravel(), rtol=5e-3, atol=2e-2)
    np.testing.assert_allclose(ce.cpu().numpy(), z["cross_loss"].ravel(), rtol=5e-3, atol=2e-2)
    # `softmax_input` (model_combine.py:138): label scores in exact fp32, the whole matrix through the bf16 operands
    S = z["softmax_input"]
    lab = z["feed_lab_input"].astype(np.int64)
    np.testing.assert_allclose(model.c_ref[:B].cpu().numpy(), S[np.arange(B), lab], rtol=1e-4, atol=1e-5)
    got = model.softmax_input(bt).cpu().numpy()
    np.testing.assert_allclose(got, S, rtol=2e-2, atol=2e-2 * np.abs(S).max())
    # top-20 of the reference-run scores (np.argsort(pred)[::-1][:20], model_combine.py:301) on margin-checked rows
    top, ngt, _ = model.eval_step(bt)
    torch.cuda.synchronize()
    order = np.argsort(-S, axis=1, kind="stable")[:, :21]
    s = np.take_along_axis(S, order, 1)
    ok = (np.abs(np.diff(s, axis=1)) > 2e-5 * np.abs(s).max(1, keepdims=True)).all(1)
    assert ok.sum() >= B // 2
    assert (top.cpu().numpy()[ok] == order[ok, :20]).all()
    rank = (S > S[np.arange(B), lab][:, None]).sum(1)
    hit = rank < 20
    assert ((ngt.cpu().numpy() < 20) == hit)[ok].all() and (ngt.cpu().numpy() == rank)[ok & hit].all()


def test_gradients_and_adam_step_match_the_reference_run():
    z, init, N = load_fixture()
    model, bt, B = build_model(z, init, N)
    model.forward_train(bt)
    model.backward(bt)
    torch.cuda.synchronize()
    got = model.ps.export_grads()
    total = np.sqrt(sum(float(z["gradnorm_" + k]) ** 2 for k in NAMES))
    bad = {}
    for k in NAMES:
        ref = grad_of(z, k).reshape(got[k].shape)
        err = np.linalg.norm(got[k].double().numpy() - ref) / (np.linalg.norm(ref) + 5e-6 * total)
        if err > 2e-2:
            bad[k] = err
    assert not bad, f"gradient vs reference run, norm-wise: {bad}"
    # neither clip norm (TensorFlow's un-aggregated reading nor the dense one) reaches max_grad = 150 here, so the
    # product's update must be the reference's
    for k in NAMES:
        assert max(float(z["clipnorm_" + k]), float(z["gradnorm_" + k])) < float(z["max_grad"])
    before = model.ps.export()
    model.apply_gradients()
    torch.cuda.synchronize()
    after = model.ps.export()
    lr = float(z["lr"])
    for k in NAMES:
        if "delta16_" + k not in z.files:
            continue
        ref = (z["delta16_" + k].astype(np.float64) * float(z["deltascale_" + k])).reshape(after[k].shape)
        d = (after[k] - before[k]).double().numpy()
        g = np.abs(grad_of(z, k).reshape(d.shape))
        firm = g > 0.05 * g.max()                  # far above the gradient error: same sign, |g| >> eps -> -lr sign(g)
        assert firm.any(), k
        np.testing.assert_allclose(d[firm], ref[firm], rtol=0, atol=0.02 * lr, err_msg=k)
