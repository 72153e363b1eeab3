"""Catalog-sharded TCAR train step: the softmax over the N candidates is split across ranks.

The data-parallel step of model_combine.Seq2SeqAttNN.train_step all-reduces the dense item gradient (364 MB at the
Globo shape) and repeats the table-wide Adam pass (2.7 GB of HBM traffic) on every rank.  Here every rank OWNS a
contiguous range of items -- the rows of the bf16 scoring operand, the fp32 master rows, their Adam moments and their
gradient -- and the step exchanges only session-sized tensors:

    all-gather   packed batch ids                 (orders the owners' previous update before the row fetch below)
    peer loads   item rows of the rank's own sessions, straight from the owners' HBM      (tcar_peer_fetch_rows)
    all-gather   Q [512,640] bf16 + label scores c [512]          -> every rank scores ALL sessions against its range
    all-reduce   softmax partial sums [R,512]                     (model_combine.py:145: CE = log sum exp(S - c))
    reduce-scat. dQ partials [R,512,640] fp32                     -> each rank keeps its own sessions' dQ
    all-gather   Qs = a_ic / sumexp [512,256] bf16                -> dense item gradient of the owned rows, complete
    all-gather   a_ic, coef, dXi of every rank's sessions         -> each owner applies the sparse rows it owns
    all-reduce   gradients of the 22 small tensors + the squared norm of the item gradient (per-tensor clip_by_norm,
                 model_combine.py:158-160, needs the norm of the WHOLE item gradient)

and then runs clip + Adam on its own rows only.  Same function as the single-GPU step (the sums over sessions and over
items are merely regrouped); verified against it by tools/dist_check.py and tests/test_gpu_parity.py.

`Seq2SeqAttNN(args)` with args["train_parallel"] == "catalog" enables it; `train_step_catalog(bt, counts)` is the
step, `sync_item_table()` re-assembles the full fp32 table + scoring operand on every rank (before evaluation, export
or a checkpoint).  With world_size == 1, args["catalog_virtual_shards"] = V makes the single process own V shards and
walk them one after the other -- the same kernels with the same shard offsets, used by the single-GPU parity test.
"""
import ctypes as C

import torch

from . import _native as nv
from . import parallel
from .params import SMALL

H, XW, KEXT, QROWS, HP = nv.H, nv.XW, nv.KEXT, nv.QROWS, nv.HP
Q_BYTES = QROWS * KEXT * 2                 # bf16 query operand
QC_BYTES = Q_BYTES + QROWS * 4             # + fp32 label scores
PAY_HEAD = QROWS * XW + QROWS              # a_ic [512,500] + coef [512] floats in front of dXi


class CatalogShardedTraining:
    """Mixin of Seq2SeqAttNN (needs its workspaces and kernels wrappers)."""

    # ------------------------------------------------------------------------------------------- set-up
    def enable_catalog_training(self, virtual_shards=1):
        dev, ps = self.dev, self.ps
        self._cat_dist = parallel.is_distributed(self.world)
        if self.world > 1 and not self._cat_dist:
            raise RuntimeError("catalog-sharded training needs an initialised torch.distributed process group")
        R = self.world if self._cat_dist else 1
        if R > nv.MAX_PEERS:
            raise ValueError(f"at most {nv.MAX_PEERS} ranks")
        self._cat_R = R
        nshards = R if self._cat_dist else max(1, int(virtual_shards))
        bounds = parallel.shard_bounds(ps.N, ps.n_pad, nshards)
        self._cat_row_bounds = parallel.catalog_row_bounds(bounds, ps.N)        # table rows, [nshards + 1]
        mine = [self.rank] if self._cat_dist else list(range(nshards))
        self._cat_shards = []
        for s in mine:
            lo, hi = bounds[s]
            n_pad = max((hi - lo + 255) // 256 * 256, 256)
            self._cat_shards.append({"lo": lo, "hi": hi, "n_pad": n_pad, "iext": ps.iext[lo: lo + n_pad],
                                     "row_lo": self._cat_row_bounds[s], "row_hi": self._cat_row_bounds[s + 1]})
        # exchange buffers; the kernels keep writing self.Q / self.c_ref / self.a_ic / self.coef / self.dXi, which now
        # alias the send buffers of the collectives (no packing kernels)
        self._qc = torch.zeros(QC_BYTES, device=dev, dtype=torch.uint8)
        self.Q = self._qc[:Q_BYTES].view(torch.bfloat16).view(QROWS, KEXT)
        self.c_ref = self._qc[Q_BYTES:].view(torch.float32)
        self._qc_all = torch.zeros(R, QC_BYTES, device=dev, dtype=torch.uint8) if R > 1 else self._qc.view(1, -1)
        self._pay = torch.zeros(PAY_HEAD + QROWS * nv.MAXT * HP, device=dev)
        self.a_ic = self._pay[: QROWS * XW].view(QROWS, XW)
        self.coef = self._pay[QROWS * XW: PAY_HEAD]
        self.dXi = self._pay[PAY_HEAD:].view(QROWS * nv.MAXT, HP)
        self._pay_all = None
        self._ids = torch.zeros(1, device=dev, dtype=torch.int32)
        self._ids_all = torch.zeros(1, device=dev, dtype=torch.int32)
        self._sumexp_all = torch.zeros(R, QROWS, device=dev)
        self._dq_all = torch.zeros(R, QROWS, KEXT, device=dev)
        self._qs_all = torch.zeros(R, QROWS, HP, device=dev, dtype=torch.bfloat16) if R > 1 else self.Qs.view(1, QROWS, HP)
        self._se_tmp, self._ce_tmp, self._dq_tmp = torch.zeros(QROWS, device=dev), torch.zeros(QROWS, device=dev), None
        if len(self._cat_shards) > 1:
            self._dq_tmp = torch.zeros(QROWS, KEXT, device=dev)
        self._sq_tmp = torch.zeros(1, device=dev)
        # tail of the small-gradient all-reduce: [theta_g | squared norm of the owned item-gradient rows]
        self._red = torch.zeros(ps.flat_size + 4, device=dev)
        for sh in self._cat_shards:
            tiles = nv.lib().tcar_score_fwd_tiles(sh["n_pad"])
            sh["tiles"] = tiles
            sh["E"] = [torch.zeros(QROWS, sh["n_pad"], device=dev, dtype=torch.bfloat16) for _ in range(R)]
            sh["part"] = [torch.zeros(tiles, QROWS, device=dev) for _ in range(R)]
            splits = max(nv.lib().tcar_score_bwd_q_splits(b, sh["n_pad"]) for b in (1, 129, 257, 385))
            sh["qpart"] = torch.zeros(splits, QROWS, KEXT, device=dev)
        # peer pointers of the fp32 item table (single node, CUDA IPC): exported once, opened once
        self._peer_ptrs = (C.c_void_p * nv.MAX_PEERS)()
        self._peer_ptrs[self.rank if self._cat_dist else 0] = ps.item_full.data_ptr()
        self._peer_bounds = (C.c_int32 * (nv.MAX_PEERS + 1))(*self._cat_row_bounds) if self._cat_dist else None
        self._peer_opened = []
        if self._cat_dist:
            self._open_peers()
        self._item_table_synced = True
        self.train_parallel = "catalog"

    def _open_peers(self):
        import torch.distributed as dist
        handle = (C.c_ubyte * nv.PEER_HANDLE_BYTES)()
        off = C.c_longlong(0)
        rc = nv.lib().tcar_peer_export(C.c_void_p(self.ps.item_full.data_ptr()), handle, C.byref(off))
        if rc != 0:
            raise nv.TcarNativeError(f"tcar_peer_export failed with code {rc} (cudaIpcGetMemHandle; the item table "
                                     "must live in a plain cudaMalloc block -- PYTORCH_CUDA_ALLOC_CONF=expandable_segments "
                                     "is not supported)")
        mine = (bytes(handle), int(off.value))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        for g, (hb, o) in enumerate(everyone):
            if g == self.rank:
                continue
            out = C.c_void_p()
            rc = nv.lib().tcar_peer_open((C.c_ubyte * nv.PEER_HANDLE_BYTES).from_buffer_copy(hb), o, C.byref(out))
            if rc != 0 or not out.value:
                raise nv.TcarNativeError(f"tcar_peer_open(rank {g}) failed with code {rc}: catalog-sharded training "
                                         "needs CUDA IPC peer access between the GPUs of the node")
            self._peer_ptrs[g] = out.value
            self._peer_opened.append((out.value, o))

    def close_peers(self):
        for ptr, off in self._peer_opened:
            nv.lib().tcar_peer_close(C.c_void_p(ptr), off)
        self._peer_opened = []

    # ------------------------------------------------------------------------------------------- the step
    def _fetch_rows(self, bt):
        """Item rows this rank's sessions read (clicks, labels, negatives) from their owners' HBM."""
        p, st = nv.ptr, nv.stream_ptr()
        for ids, n, add in ((bt.seq, bt.B * bt.T, 0), (bt.label, bt.B, 1), (bt.neg, bt.B * bt.Nn if bt.Nn else 0, 1)):
            if n == 0:
                continue
            nv.LAUNCHES["count"] += 1
            rc = nv.lib().tcar_peer_fetch_rows(p(ids), n, add, self._peer_ptrs, self._peer_bounds, self._cat_R,
                                               self.rank, p(self.ps.item_full), st)
            if rc != 0:
                raise nv.TcarNativeError(f"tcar_peer_fetch_rows failed with code {rc}")

    def _gather(self, out_flat, inp_flat):
        """all_gather_into_tensor on flat views: out = [rank 0 | rank 1 | ...]."""
        import torch.distributed as dist
        dist.all_gather_into_tensor(out_flat, inp_flat)

    def train_step_catalog(self, bt, counts=None):
        """One train step with the catalog sharded across ranks.  bt = this rank's sessions (B may be 0 for a tail
        batch); counts[g] = sessions of rank g in this step (default: bt.B on every rank) -- all ranks must pass the
        same list, and the same T / Nn.  Returns this rank's loss [B]."""
        import torch.distributed as dist
        ps, p = self.ps, nv.ptr
        R, me, on = self._cat_R, (self.rank if self._cat_dist else 0), self._cat_dist
        B, T, Nn = bt.B, bt.T, bt.Nn
        counts = [B] * R if counts is None else [int(c) for c in counts]
        if len(counts) != R or counts[me] != B:
            raise ValueError("counts must list the sessions of every rank (and counts[rank] == bt.B)")
        Bmax = max(counts)
        if Bmax == 0:
            return self.loss[:0]
        self.sync_updates()
        self._prefetched = None
        self._item_table_synced = False
        groups = [g for g in range(R) if counts[g] > 0]
        L = 7 * Bmax * T + 3 * Bmax + Bmax * Nn
        # ---- packed ids of every rank (sparse-row scatter below); as the first collective of the step it also orders
        # every owner's previous Adam pass before the peer loads
        if on:
            if self._ids.numel() < L:
                self._ids = torch.zeros(L, device=self.dev, dtype=torch.int32)
                self._ids_all = torch.zeros(R * L, device=self.dev, dtype=torch.int32)
            if B > 0:
                self._ids[: bt.buf.numel()].copy_(bt.buf)
            self._gather(self._ids_all[: R * L], self._ids[:L])
            ids_of = lambda g: self._ids_all[g * L: (g + 1) * L]
            if B > 0:
                self._fetch_rows(bt)
        else:
            ids_of = lambda g: bt.buf
        # ---- session forward of the local sessions -> Q, c_ref (views of the send buffer)
        if B > 0:
            self._session_forward(bt)
        if on:
            self._gather(self._qc_all.view(-1), self._qc)
        qc = self._qc_all
        q_of = lambda g: qc[g, :Q_BYTES]
        c_of = lambda g: qc[g, Q_BYTES:]
        # ---- every session group against every owned item range
        multi = len(self._cat_shards) > 1
        if multi:
            self._sumexp_all.zero_()
        for sh in self._cat_shards:
            if sh["hi"] <= sh["lo"]:
                continue
            for g in groups:
                nv.counted_call("tcar_score_fwd", 1, p(q_of(g)), p(sh["iext"]), p(c_of(g)), p(sh["E"][g]),
                                p(sh["part"][g]), None, None, counts[g], sh["hi"] - sh["lo"], sh["n_pad"], 0,
                                self._cluster_for(counts[g]))
                dst = self._se_tmp if multi else self._sumexp_all[g]
                nv.counted_call("tcar_ce_finish", 1, p(sh["part"][g]), p(dst), p(self._ce_tmp), sh["tiles"], counts[g])
                if multi:
                    self._sumexp_all[g, : counts[g]] += self._se_tmp[: counts[g]]
        if on:
            dist.all_reduce(self._sumexp_all, op=dist.ReduceOp.SUM)
        if B > 0:
            self.sumexp[:B].copy_(self._sumexp_all[me, :B])
            torch.log(self.sumexp[:B], out=self.ce[:B])
            nv.counted_call("tcar_neg_loss", 1, p(self.a_ic), p(ps.item), p(ps.content), p(bt.neg), p(self.ce),
                            p(self.negloss), p(self.loss), p(self.coef), p(self.dA_neg), B, Nn)
        # ---- dQ: partial sums over the owned items for every group, reduce-scattered to the sessions' ranks
        first = True
        for sh in self._cat_shards:
            if sh["hi"] <= sh["lo"]:
                continue
            for g in groups:
                dst = self._dq_all[g] if first else self._dq_tmp
                nv.counted_call("tcar_score_bwd_q", 2, p(sh["E"][g]), p(sh["iext"]), p(sh["qpart"]), p(dst), counts[g],
                                sh["n_pad"])
                if not first:
                    rows = (counts[g] + 127) // 128 * 128
                    self._dq_all[g, :rows] += self._dq_tmp[:rows]
            first = False
        if on:
            dist.reduce_scatter_tensor(self.dq_raw, self._dq_all, op=dist.ReduceOp.SUM)
            dq_raw = self.dq_raw
        else:
            dq_raw = self._dq_all[0]
        if B > 0:
            nv.counted_call("tcar_score_bwd_finish", 1, p(dq_raw), p(self.sumexp), p(self.dA_neg), p(self.a_ic),
                            p(ps.ct_tab), p(ps.item), p(ps.content), p(ps.mwdhm), p(bt.label), p(self.d_a_ic),
                            p(self.d_a_pt), p(self.dTq), p(self.Qs), B)
        if on:
            self._gather(self._qs_all.view(-1), self.Qs.view(-1))
        # ---- dense gradient of the owned item rows: complete after the last group (sum over ALL sessions)
        for sh in self._cat_shards:
            if sh["hi"] <= sh["lo"]:
                continue
            g_rows = ps.item_g_full[sh["lo"]:]            # the kernel writes table row n + 1 for local item n
            for k, g in enumerate(groups):
                nv.counted_call("tcar_score_bwd_i_acc", 1, p(sh["E"][g]), p(self._qs_all[g]), p(g_rows), None,
                                counts[g], sh["hi"] - sh["lo"], sh["n_pad"], 1 if k > 0 else 0)
        # ---- session-side backward of the local sessions, then the sparse rows of every rank's sessions
        if B > 0:
            self._session_backward(bt)
        else:
            ps.theta_g.zero_()
        n_pay = PAY_HEAD + Bmax * T * HP
        if on:
            if self._pay_all is None or self._pay_all.numel() < R * n_pay:
                self._pay_all = torch.zeros(R * n_pay, device=self.dev)
            self._gather(self._pay_all[: R * n_pay], self._pay[:n_pay])
            pay_of = lambda g: self._pay_all[g * n_pay: (g + 1) * n_pay]
        else:
            pay_of = lambda g: self._pay
        self._alloc_scatter(Bmax * T + Bmax + Bmax * Nn)
        for sh in self._cat_shards:
            if sh["row_hi"] <= sh["row_lo"]:
                continue
            for g in groups:
                Bg, ids, pay = counts[g], ids_of(g), pay_of(g)
                Mg = Bg * T
                seq, label = ids[:Mg], ids[7 * Mg + 2 * Bg: 7 * Mg + 3 * Bg]
                neg = ids[7 * Mg + 3 * Bg: 7 * Mg + 3 * Bg + Bg * Nn] if Nn else None
                nv.counted_call("tcar_scatter_add_rows_range", 3, p(seq), p(label), p(neg), p(pay[PAY_HEAD:]),
                                p(pay[: QROWS * XW]), p(pay[QROWS * XW: PAY_HEAD]), p(ps.item), p(ps.item_g),
                                p(self.hash_keys), p(self.hash_cnt), p(self.hash_acc), p(self.entry_slot), None,
                                self.hash_size, Bg, T, Nn, sh["row_lo"], sh["row_hi"])
        # ---- clip norms: small tensors after their all-reduce; the item tensor from the owned rows, summed over ranks
        red = self._red
        red[: ps.flat_size].copy_(ps.theta_g)
        red[ps.flat_size:].zero_()
        for sh in self._cat_shards:
            if sh["row_hi"] <= sh["row_lo"]:
                continue
            rows = ps.item_g_full[sh["row_lo"]: sh["row_hi"]]
            nv.counted_call("tcar_sqnorm_big", 2, p(rows), p(ps.norm_partial), p(self._sq_tmp), rows.numel())
            red[ps.flat_size: ps.flat_size + 1] += self._sq_tmp
        if on:
            dist.all_reduce(red, op=dist.ReduceOp.SUM)
        ps.theta_g.copy_(red[: ps.flat_size])
        ps.sqnorm_item.copy_(red[ps.flat_size: ps.flat_size + 1])
        nv.counted_call("tcar_sqnorm_segments", 1, p(ps.theta_g), p(ps.seg_off), p(ps.sqnorm_small), len(SMALL))
        ps.step.add_(1)
        self.global_step += 1
        self._update_small()
        for sh in self._cat_shards:
            lo, n = sh["row_lo"], sh["row_hi"] - sh["row_lo"]
            if n <= 0:
                continue
            nv.counted_call("tcar_adam_item", 1, p(ps.item_full[lo:]), p(ps.item_m_full[lo:]), p(ps.item_v_full[lo:]),
                            p(ps.item_g_full[lo:]), p(ps.sqnorm_item), p(ps.step), self.lr, self.max_grad_f,
                            p(ps.iext), lo, n, None, 0)
        self._fused_norm = False
        return self.loss[:B]

    # ------------------------------------------------------------------------------------------- re-assembly
    def sync_item_table(self):
        """Every rank receives every owner's fp32 rows and rebuilds the whole bf16 scoring operand: evaluation, export
        and checkpoints read the full table.  A no-op when nothing was trained since the last call."""
        if getattr(self, "train_parallel", "dp") != "catalog" or self._item_table_synced:
            return
        self.sync_updates()
        if self._cat_dist:
            import torch.distributed as dist
            rb = self._cat_row_bounds
            for g in range(self._cat_R):
                if rb[g + 1] > rb[g]:
                    for t in (self.ps.item_full, self.ps.item_m_full, self.ps.item_v_full):
                        dist.broadcast(t[rb[g]: rb[g + 1]], src=g)
        self.ps.rebuild_iext()
        self._item_table_synced = True
