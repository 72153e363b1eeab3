"""GPU-resident sampler.  Same constructor, shuffles and batch order as `sampler.Sampler` (which mirrors
the reference's sampler.py:24-50), but the per-batch gather of the 7 index planes, the click context and the labels
(sampler.py:52-113) runs on the device from a columnar cache uploaded once per split (`tcar_assemble_batch`): the
host sends only the batch's B bucket rows -- plus, in `negatives="host"` mode, the B x Nn negatives it drew from the
reference's NumPy stream (np.random.randint, sampler.py:98-99), which makes every batch bit-identical to
`Sampler.next_packed()`.  `negatives="device"` draws them on the device (Philox4x32-10 keyed by `seed`; restated in
oracle/philox_oracle.py): 2 KB of host->device traffic per batch instead of 83 KB, no per-batch NumPy work.  With
`negative_mode="impression"` (the MIND configuration, sampler.py:96,118-131) the device mode runs the reference's
impression-list algorithm in `tcar_impression_negatives` from a CSR copy of the sessions' impression lists, already
mapped through item_dict; the host mode keeps Python's `random.choice` stream and stays bit-identical to the reference."""
import numpy as np
import torch

from . import _native as nv
from .sampler import Sampler, _columnar

_DEVICE_CACHE = {}


class _DeviceColumnar:
    def __init__(self, col, device):
        self.seq = {L: torch.from_numpy(np.ascontiguousarray(a)).to(device) for L, a in col.seq.items()}
        self.feats = {L: torch.from_numpy(np.ascontiguousarray(a)).to(device) for L, a in col.feats.items()}
        self.ctx = {L: torch.from_numpy(np.ascontiguousarray(a)).to(device) for L, a in col.ctx.items()}


def _device_columnar(col, device):
    key = (id(col), str(device))
    ent = _DEVICE_CACHE.get(key)
    if ent is None or ent[0] is not col:
        ent = _DEVICE_CACHE[key] = (col, _DeviceColumnar(col, device))
    return ent[1]


def impression_csr(col, session_dict, neighbor_dict, item_dict):
    """Per session-length bucket L: CSR (offsets [n + 1], ids) of every session's impression list in bucket-row order,
    mapped through item_dict to 0-based item ids (-1 = article not in item_dict; sampler.py:119,124-125).  The list
    of session key "<sid>_<len>" is neighbor_dict[int(sid)] (sampler.py:96); a session without a list gets an empty
    one (all its negatives are then uniform fills)."""
    keys = {}
    for key, r in col.row.items():
        keys.setdefault(len(session_dict[key]) - 1, {})[r] = key
    out = {}
    for L, rows in keys.items():
        off = np.zeros(len(rows) + 1, dtype=np.int32)
        ids = []
        for r in range(len(rows)):
            lst = neighbor_dict.get(int(str(rows[r]).split("_")[0]), ())
            ids.extend(item_dict.get(x, 0) - 1 for x in lst)
            off[r + 1] = len(ids)
        out[L] = (off, np.asarray(ids if ids else [-1], dtype=np.int32))
    return out


def _device_impressions(col, session_dict, neighbor_dict, item_dict, device):
    key = (id(col), id(neighbor_dict), id(item_dict), str(device))
    ent = _DEVICE_CACHE.get(key)
    if ent is None or ent[0] is not col or ent[2] is not neighbor_dict or ent[3] is not item_dict:
        csr = impression_csr(col, session_dict, neighbor_dict, item_dict)
        ent = _DEVICE_CACHE[key] = (col, {L: (torch.from_numpy(o).to(device), torch.from_numpy(i).to(device))
                                          for L, (o, i) in csr.items()}, neighbor_dict, item_dict)
    return ent[1]


class DeviceSampler(Sampler):
    def __init__(self, model, len_dict, session_dict, session_time_dict=None, neighbor_dict=None, item_dict=None,
                 neg_num=None, batch_size=1024, negative_mode="uniform", negatives="host", seed=2020, rank=0, world=1,
                 verbose=True):
        super().__init__(len_dict, session_dict, session_time_dict, neighbor_dict, item_dict, neg_num, batch_size,
                         negative_mode, verbose)
        if negatives not in ("host", "device"):
            raise ValueError("negatives must be 'host' (reference NumPy stream) or 'device' (Philox)")
        self.model, self.negatives, self.seed = model, negatives, int(seed)
        self.rank, self.world = rank, world
        self._col = _columnar(session_dict, session_time_dict)
        self._dev = _device_columnar(self._col, model.dev)
        self._counter = 0                     # Philox counter blocks consumed so far (device negatives)
        self._impr = None
        if negatives == "device" and negative_mode == "impression" and neighbor_dict and neg_num:
            self._impr = _device_impressions(self._col, session_dict, neighbor_dict, item_dict, model.dev)

    def host_part(self):
        """Bucket rows (and host-drawn negatives) of the next batch: (small int32 array, B, T, Nn).  Consumes the
        NumPy stream exactly like Sampler.next_packed()."""
        ids = self.session_id_batches[self.batch_i]
        B = len(ids)
        T = len(self.session_dict[ids[0]]) - 1
        Nn = self.neg_num if (self.neighbor_dict and self.neg_num) else 0
        rows = np.fromiter((self._col.row[k] for k in ids), dtype=np.int32, count=B)
        negs = None
        if Nn and self.negatives == "host":
            if self.negative_mode == "impression":
                negs = np.array([self._negatives(sid) for sid in ids], dtype=np.int32).reshape(B, Nn)
            else:
                negs = np.random.randint(0, self.item_num, size=(B, Nn)).astype(np.int32)
        self.batch_i += 1
        if self.world > 1:
            # data parallel: every rank walks the same batches and keeps its slice of the sessions
            from .parallel import shard_sessions
            lo, hi = shard_sessions(B, self.rank, self.world)
            first_neg = lo * Nn                  # element index of this slice inside the global batch's negatives
            rows, negs, B_glob, B = rows[lo:hi], (negs[lo:hi] if negs is not None else None), B, hi - lo
        else:
            first_neg, B_glob = 0, B
        small = rows if negs is None else np.concatenate([rows, negs.reshape(-1)])
        return np.ascontiguousarray(small, dtype=np.int32), B, T, Nn, B_glob, first_neg

    def next_device(self):
        """The next batch, assembled on the device -> model_combine.Batch."""
        from .model_combine import Batch
        small, B, T, Nn, B_glob, first_neg = self.host_part()
        model, p = self.model, nv.ptr
        total = 7 * B * T + 3 * B + B * Nn
        out = torch.empty(max(total, 1), device=model.dev, dtype=torch.int32)
        offset = self._counter
        impr_blocks = (21 + Nn + 3) // 4                 # TCAR_IMPR_BLOCKS(Nn)
        if self.negatives == "device" and Nn:
            self._counter += B_glob * impr_blocks if self._impr is not None else (B_glob * Nn + 3) // 4
        counts = None
        if self.world > 1:
            from .parallel import catalog_counts
            counts = catalog_counts(B_glob, self.world)
        if B == 0:
            bt = Batch(out[:0], 0, T, Nn)
            bt.counts = counts
            return bt
        host = model.stage_to_device(small, 0, 0, 0).buf          # pinned ring -> device, one copy
        rows_d = host[:B]
        neg_d = host[B:] if (Nn and self.negatives == "host") else None
        dev = self._dev
        if self._impr is not None and Nn:
            # impression-list negatives (sampler.py:118-131) drawn on the device, then handed to the assembly kernel
            off_d, ids_d = self._impr[T]
            neg_d = torch.empty(B * Nn, device=model.dev, dtype=torch.int32)
            nv.counted_call("tcar_impression_negatives", 1, p(rows_d), p(off_d), p(ids_d), B, Nn, int(self.item_num),
                            self.seed, offset + (first_neg // Nn) * impr_blocks, p(neg_d))
        elif first_neg % 4:
            raise ValueError("device negatives need shard boundaries on 4-element Philox blocks (B * Nn % 4 == 0)")
        nv.counted_call("tcar_assemble_batch", 1, p(rows_d), p(dev.seq[T]), p(dev.feats[T]), p(dev.ctx[T]),
                        int(dev.seq[T].shape[0]), B, T, Nn, p(neg_d), int(getattr(self, "item_num", 0) or 0),
                        self.seed, offset + first_neg // 4, p(out))
        bt = Batch(out[:total], B, T, Nn)
        bt.counts = counts
        return bt
