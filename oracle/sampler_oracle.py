"""CPU ORACLE (test infrastructure only) for the reference's host-side index production, sampler.py:18-140.

Restates, list-for-list, what `Sampler.__init__` / `next_batch` / `neg_neighbor_from_impre` compute, consuming the
same two global RNG streams in the same order (python `random` for the shuffles and impression picks, the legacy
NumPy global stream for uniform negatives).  Pinned against the reference's own sampler.py run in the build
container: tests/golden/sampler_ref.json (tests/test_oracle_golden.py).
"""
import math
import random

import numpy as np


def bucketized(seconds):
    """sampler.py:18-21: searchsorted([0..10], log2(seconds+1)) (side='left').  Returns 0..11; 11 (>= 1024 s) is
    out of range for the 11-row duration table -- callers clamp (SURVEY gotcha 7)."""
    t = math.log2(seconds + 1)
    b = 0
    while b < 11 and b < t:      # first boundary index with boundaries[b] >= t
        b += 1
    return b


class SamplerOracle:
    def __init__(self, len_dict, session_dict, session_time_dict=None, neighbor_dict=None, item_dict=None,
                 neg_num=None, batch_size=1024):
        self.batch_size = batch_size
        self.neighbor_dict, self.item_dict, self.neg_num = neighbor_dict, item_dict, neg_num
        if item_dict is not None:
            self.item_num = len(item_dict)
        self.session_dict, self.session_time_dict = session_dict, session_time_dict
        self.batches = []
        for _slen, ids in len_dict.items():            # sampler.py:40-48 (in-place shuffle of the caller's lists)
            random.shuffle(ids)
            while len(ids) > batch_size:               # strict '>' : an exact multiple keeps a full last chunk
                self.batches.append(ids[:batch_size])
                ids = ids[batch_size:]
            if len(ids):
                self.batches.append(ids)
        random.shuffle(self.batches)                   # sampler.py:49
        self.i = 0

    def has_next(self):
        return self.i < len(self.batches)

    def next_batch(self):
        b_in, b_out, neg_all, gap_all = [], [], [], []
        pt = [[], [], [], [], []]                      # month, day, week, hour, minute
        ct = [[], [], [], [], []]                      # month, day, week, hour, minute of the LAST input click
        for sid in self.batches[self.i]:
            seq = self.session_dict[sid]
            b_in.append(seq[:-1])
            b_out.append(seq[-1] - 1)
            neg, gap = [], []
            if self.session_time_dict:
                cols = [[], [], [], [], []]
                last_click = None
                for t in self.session_time_dict[sid][:-1]:
                    p = t["publish_t"]
                    for c, v in zip(cols, (p.month, p.day, p.isoweekday(), p.hour + 1, p.minute + 1)):
                        c.append(v)                    # sampler.py:80-85
                    last_click = t["click_t"]
                    gap.append(bucketized(t["active_t"]))          # sampler.py:87
                if self.neighbor_dict:
                    while len(neg) < self.neg_num:                 # sampler.py:98-99
                        neg.append(int(np.random.randint(0, self.item_num)))
                for dst, c in zip(pt, cols):
                    dst.append(c)
                c = last_click                                      # sampler.py:105-109
                for dst, v in zip(ct, (c.month - 1, c.day - 1, c.isoweekday() - 1, c.hour, c.minute)):
                    dst.append(v)
            neg_all.append(neg)
            gap_all.append(gap)
        self.i += 1
        return b_in, b_out, tuple(pt), tuple(ct), neg_all, gap_all

    def neg_neighbor_from_impre(self, sessionid):
        """sampler.py:118-131: up to 21 random.choice tries over the impression list, then uniform fill."""
        cand = self.neighbor_dict[sessionid]
        neg, cnt = [], 0
        while len(neg) < self.neg_num:
            cnt += 1
            pick = random.choice(cand)
            if pick in self.item_dict:
                neg.append(self.item_dict[pick] - 1)
            if cnt > 20:
                break
        while len(neg) < self.neg_num:
            neg.append(int(np.random.randint(0, self.item_num)))
        return neg
