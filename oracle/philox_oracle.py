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
