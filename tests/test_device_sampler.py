"""GPU-resident sampler: Philox restatement pinned on the published Random123 known-answer vectors (CPU),
device-assembled batches bit-identical to Sampler.next_packed() (GPU), device negatives bit-identical to the oracle."""
import random

import numpy as np
import pytest
import torch

from oracle import philox_oracle as PO


def test_philox_known_answer_vectors():
    # Random123 kat_vectors: philox4x32_10, counter = key = 0, and counter = key = all ones (counter words 2, 3 are
    # fixed at zero in this restatement, so only the first vector applies verbatim)
    got = PO.philox4x32_10(0, np.array([0], dtype=np.uint64))[0]
    assert [int(x) for x in got] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_device_negatives_are_uniform_and_reproducible():
    a = PO.device_negatives(2020, 7, 512 * 20, 364047)
    b = PO.device_negatives(2020, 7, 512 * 20, 364047)
    assert np.array_equal(a, b)
    assert a.min() >= 0 and a.max() < 364047
    # a different counter offset gives a different block of the same stream: shifting by one 4-element block aligns
    c = PO.device_negatives(2020, 8, 512 * 20 - 4, 364047)
    assert np.array_equal(a[4:], c)
    assert abs(a.mean() / 364047 - 0.5) < 0.02


def _split(N=3000, n_sessions=3000, seed=5):
    from tcar_b200 import synth
    return synth.make_sessions(N, n_sessions, seed=seed)


def _model(N=3000):
    from tcar_b200 import synth
    from tcar_b200.model_combine import Seq2SeqAttNN
    content, mwdhm, category = synth.make_catalog(N, seed=3)
    np.random.seed(2020)
    return Seq2SeqAttNN(dict(publish_time_MWDHM=mwdhm, itemnum=N, category_id=None, item_freq_dict_norm={},
                             reverse_item=None, content_emb=content, emb_stddev=0.002, stddev=0.05, hidden_size=250,
                             time_hidden_size=64, l2_emb=0.0, batch_size=128, epoch=1, neg_num=20, lr=0.001,
                             max_grad=150))


@pytest.mark.gpu
def test_device_assembled_batches_equal_host_batches():
    from tcar_b200.device_sampler import DeviceSampler
    from tcar_b200.sampler import Sampler
    ld, sd, td, idict, impr = _split()
    model = _model()
    import copy
    # both samplers consume the GLOBAL NumPy stream (negatives), so run them one after the other from the same seeds
    random.seed(7); np.random.seed(7)
    host = Sampler(copy.deepcopy(ld), sd, td, impr, idict, 20, batch_size=128, verbose=False)
    want = []
    while host.has_next():
        want.append(host.next_packed())
    random.seed(7); np.random.seed(7)
    dev = DeviceSampler(model, copy.deepcopy(ld), sd, td, impr, idict, 20, batch_size=128, negatives="host",
                        verbose=False)
    for n, (packed, B, T, Nn) in enumerate(want):
        bt = dev.next_device()
        assert (bt.B, bt.T, bt.Nn) == (B, T, Nn)
        assert np.array_equal(bt.buf.cpu().numpy(), packed), f"batch {n}"
    assert not dev.has_next() and len(want) > 10


@pytest.mark.gpu
def test_device_negatives_match_the_philox_oracle_and_shard_consistently():
    from tcar_b200.device_sampler import DeviceSampler
    ld, sd, td, idict, impr = _split()
    model = _model()
    import copy
    random.seed(9)
    full = DeviceSampler(model, copy.deepcopy(ld), sd, td, impr, idict, 20, batch_size=128, negatives="device",
                         seed=11, verbose=False)
    random.seed(9)
    half = DeviceSampler(model, copy.deepcopy(ld), sd, td, impr, idict, 20, batch_size=128, negatives="device",
                         seed=11, rank=1, world=2, verbose=False)
    offset = 0
    for _ in range(6):
        bt = full.next_device()
        neg = bt.neg.cpu().numpy()
        want = PO.device_negatives(11, offset, bt.B * 20, full.item_num)
        assert np.array_equal(neg, want)
        offset += (bt.B * 20 + 3) // 4
        hb = half.next_device()
        lo = bt.B - hb.B if bt.B > 1 else 0
        from tcar_b200.parallel import shard_sessions
        lo, hi = shard_sessions(bt.B, 1, 2)
        assert np.array_equal(hb.neg.cpu().numpy(), neg.reshape(bt.B, 20)[lo:hi].reshape(-1))
        assert np.array_equal(hb.seq.cpu().numpy().reshape(hb.B, hb.T), bt.seq.cpu().numpy().reshape(bt.B, bt.T)[lo:hi])


@pytest.mark.gpu
def test_training_with_device_sampler_matches_host_sampler():
    """Same losses, step by step, whether batches are packed on the host or assembled on the device."""
    from tcar_b200.device_sampler import DeviceSampler
    from tcar_b200.sampler import Sampler
    import copy
    ld, sd, td, idict, impr = _split(n_sessions=1200)
    losses = []
    for use_dev in (False, True):
        model = _model()
        random.seed(3); np.random.seed(3)
        if use_dev:
            s = DeviceSampler(model, copy.deepcopy(ld), sd, td, impr, idict, 20, batch_size=128, negatives="host",
                              verbose=False)
        else:
            s = Sampler(copy.deepcopy(ld), sd, td, impr, idict, 20, batch_size=128, verbose=False)
        out = []
        for _ in range(5):
            bt = s.next_device() if use_dev else model.stage_to_device(*s.next_packed())
            out.append(model.train_step(bt).clone())
        torch.cuda.synchronize()
        losses.append(torch.cat(out))
    assert torch.equal(losses[0], losses[1])


# ------------------------------------------------------------------------------------------- impression-list negatives
def test_impression_oracle_is_the_reference_algorithm_on_another_stream(monkeypatch):
    """oracle.philox_oracle.device_impression_negatives restates sampler.py:118-131 with Philox words in place of
    random.choice / np.random.randint.  Drive the product's host implementation (pinned on the reference's own output
    by test_host_logic / sampler_ref.json) with the SAME words: the two must produce the same negatives."""
    from tcar_b200 import sampler as S, synth
    from tcar_b200.device_sampler import impression_csr
    N, Nn, seed, offset = 500, 12, 99, 5
    ld, sd, td, idict, _ = synth.make_sessions(N, 300, seed=4)
    impr = synth.make_impressions(N, 300, seed=4, mean_len=6.0, miss=0.6)        # many misses: tries run out
    impr[3] = ["x1", "x2"]                                                        # nothing usable: all fills
    impr[5] = ["a7"]                                                              # a single usable article
    smp = S.Sampler({k: list(v) for k, v in ld.items()}, sd, td, impr, idict, Nn, batch_size=64, verbose=False)
    col = S._columnar(sd, td)
    csr = impression_csr(col, sd, impr, idict)
    nb = PO.impr_blocks(Nn)
    for L in sorted(csr)[:4]:
        keys = [k for k in sd if len(sd[k]) - 1 == L][:40]
        rows = np.array([col.row[k] for k in keys])
        want = PO.device_impression_negatives(seed, offset, rows, csr[L][0], csr[L][1], Nn, smp.item_num)
        for b, key in enumerate(keys):
            words = PO.philox4x32_10(seed, np.uint64(offset + b * nb) + np.arange(nb, dtype=np.uint64)).reshape(-1)
            st = {"tries": 0, "found": 0, "fills": 0}

            def choice(seq):
                pick = seq[(int(words[st["tries"]]) * len(seq)) >> 32]
                st["tries"] += 1
                st["found"] += pick in idict
                return pick

            def randint(lo, hi):
                v = (int(words[21 + st["found"] + st["fills"]]) * hi) >> 32
                st["fills"] += 1
                return v

            monkeypatch.setattr(S.random, "choice", choice)
            monkeypatch.setattr(S.np.random, "randint", randint)
            got = smp.neg_neighbor_from_impre(int(key.split("_")[0]))
            assert got == want[b].tolist(), (L, b, key)
            assert st["tries"] <= 21


@pytest.mark.gpu
def test_device_impression_negatives_match_the_oracle_and_shard_consistently():
    from tcar_b200 import synth
    from tcar_b200.device_sampler import DeviceSampler, impression_csr
    from tcar_b200.sampler import Sampler, _columnar
    import copy
    N, Nn = 3000, 50
    ld, sd, td, idict, impr = synth.make_sessions(N, 2500, seed=6, impressions="mind")
    impr[0] = ["x1"]                              # a session whose list holds no known article
    model = _model(N)
    random.seed(9)
    full = DeviceSampler(model, copy.deepcopy(ld), sd, td, impr, idict, Nn, batch_size=128, negative_mode="impression",
                         negatives="device", seed=13, verbose=False)
    random.seed(9)
    half = DeviceSampler(model, copy.deepcopy(ld), sd, td, impr, idict, Nn, batch_size=128, negative_mode="impression",
                         negatives="device", seed=13, rank=1, world=2, verbose=False)
    random.seed(9)
    host = Sampler(copy.deepcopy(ld), sd, td, impr, idict, Nn, batch_size=128, negative_mode="impression", verbose=False)
    col = _columnar(sd, td)
    csr = impression_csr(col, sd, impr, idict)
    offset, in_list = 0, 0
    for _ in range(8):
        ids = full.session_id_batches[full.batch_i]
        bt = full.next_device()
        packed, B, T, _ = host.next_packed()
        assert (bt.B, bt.T, bt.Nn) == (B, T, Nn)
        assert np.array_equal(bt.buf[: 7 * B * T + 3 * B].cpu().numpy(), packed[: 7 * B * T + 3 * B])
        rows = np.array([col.row[k] for k in ids])
        want = PO.device_impression_negatives(13, offset, rows, csr[T][0], csr[T][1], Nn, full.item_num)
        got = bt.neg.cpu().numpy().reshape(B, Nn)
        assert np.array_equal(got, want)
        offset += B * PO.impr_blocks(Nn)
        for b, k in enumerate(ids):
            allowed = {idict[x] - 1 for x in impr.get(int(k.split("_")[0]), ()) if x in idict}
            in_list += sum(int(v) in allowed for v in got[b][:3])
        hb = half.next_device()
        from tcar_b200.parallel import shard_sessions
        lo, hi = shard_sessions(B, 1, 2)
        assert np.array_equal(hb.neg.cpu().numpy().reshape(-1, Nn), got[lo:hi])
    assert in_list > 0
