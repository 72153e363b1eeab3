"""TEST INFRASTRUCTURE ONLY (oracle/): NumPy restatement of the counter-based generator behind the device-side
negative sampler (tcar_assemble_batch with neg_in == NULL).  Philox4x32-10 (Salmon et al., SC'11): key = (seed lo,
seed hi), counter = (ctr lo, ctr hi, 0, 0); negative e of a batch = (word[e % 4] of counter offset + e // 4) mapped to
[0, item_num) by (x * item_num) >> 32.  The reference draws its negatives with np.random.randint (sampler.py:98-99);
that host stream stays available (bit-exact mode) -- this generator is the throughput mode of SURVEY 8f-2."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(seed, counters):
    """counters: uint64 array [n] -> uint32 array [n, 4]."""
    ctr = np.asarray(counters, dtype=np.uint64)
    c0, c1 = ctr & MASK, ctr >> np.uint64(32)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2                       # 64-bit products of 32-bit values
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def device_negatives(seed, offset, count, item_num):
    """The `count` negatives tcar_assemble_batch draws for one batch (row-major [B, Nn] flattened)."""
    e = np.arange(count, dtype=np.uint64)
    words = philox4x32_10(seed, np.uint64(offset) + (e >> np.uint64(2)))
    x = words[np.arange(count), (e & np.uint64(3)).astype(np.int64)].astype(np.uint64)
    return ((x * np.uint64(item_num)) >> np.uint64(32)).astype(np.int32)


def impr_blocks(neg_num):
    """TCAR_IMPR_BLOCKS: Philox counters one session owns in tcar_impression_negatives."""
    return (21 + neg_num + 3) // 4


def device_impression_negatives(seed, offset, rows, impr_off, impr_ids, neg_num, item_num):
    """Restatement of tcar_impression_negatives = sampler.py:118-131 (neg_neighbor_from_impre) driven by the Philox
    stream: <= 21 draws `random.choice(neighbor_set)`, kept when the article is in item_dict (impr_ids >= 0, 0-based
    item id), until neg_num are found; then `np.random.randint(0, item_num)` fills the rest.  Returns [B, neg_num]."""
    B = len(rows)
    nb = impr_blocks(neg_num)
    out = np.zeros((B, neg_num), dtype=np.int32)
    for b, r in enumerate(rows):
        words = philox4x32_10(seed, np.uint64(offset) + np.uint64(b * nb) + np.arange(nb, dtype=np.uint64)).reshape(-1)
        lo, ln = int(impr_off[r]), int(impr_off[r + 1] - impr_off[r])
        neg, cnt = [], 0
        if ln > 0:
            while len(neg) < neg_num:                       # sampler.py:121
                cnt += 1
                pick = int(impr_ids[lo + ((int(words[cnt - 1]) * ln) >> 32)])      # random.choice(neighor_set)
                if pick >= 0:                               # `if randomid in self.item_dict`
                    neg.append(pick)
                if cnt > 20:                                # sampler.py:126-127
                    break
        while len(neg) < neg_num:                           # sampler.py:128-129
            neg.append((int(words[21 + len(neg)]) * item_num) >> 32)
        out[b] = neg
    return out
