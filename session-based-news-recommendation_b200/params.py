"""Device-resident parameter store of the TCAR model (layout chosen for B200).

  item / content      [N+1, 256] fp32   256-float row pitch -> 128-bit aligned gathers (cols 250..255 are zero)
  item Adam m, v, g   [N+1, 256] fp32
  iext                [Npad, 640] bf16  scoring operand [item | content | one-hot time bins | 0]
  theta / m / v / g   flat fp32         the 22 small tensors, concatenated in tf.trainable_variables() order
                                        (model_combine.py:151) with a segment table for per-tensor clip_by_norm
"""
import numpy as np
import torch

from . import _native as nv

H, HP, TH = nv.H, nv.HP, nv.TH

# tf.trainable_variables() order of model_combine.py (item first, then these 22)
SMALL = [
    ("pos", (40, H)), ("month", (13, TH)), ("day", (32, TH)), ("week", (8, TH)), ("hour", (25, TH)),
    ("minute", (61, TH)), ("dur", (11, TH)),
    ("W_in", (2 * H, H)), ("W_c", (H, H)), ("W_i", (TH, H)), ("w_r", (H, 1)),
    ("Wq1", (2 * TH, H)), ("bq1", (H,)), ("Wq2", (H, 2 * H)), ("bq2", (2 * H,)),
    ("W_a", (2 * H, 2 * H)), ("b_a", (2 * H,)),
    ("W1", (5 * TH, H)), ("W2", (H, H)), ("w_t", (H, 1)),
    ("W_p", (5 * TH, 5 * TH)), ("b_p", (5 * TH,)),
]
PARAM_ORDER = ["item"] + [n for n, _ in SMALL]

# Pre-split (tf32 hi / lo), padded copies of the dense weights that feed tcar_gemm_tf32:
#   name -> (source tensor, rows of the padded tensor, first destination row, pitch, folded second source)
# "W_in1" = W_in + [0; W_c]   (X.W_in + Xc.W_c = X.W_in1: Xc = X[:, 250:] is not 16-byte aligned for TMA)
# "W2x"   = [0; W2]           (Xc.W2 = X.W2x for the same reason)
PREPPED = [
    ("W_in1", "W_in", 2 * H, 0, 256, "W_c"), ("W_i", "W_i", TH, 0, 256, None), ("W1", "W1", 5 * TH, 0, 256, None),
    ("W2x", "W2", 2 * H, H, 256, None), ("Wq1", "Wq1", 2 * TH, 0, 256, None), ("Wq2", "Wq2", H, 0, 512, None),
    ("W_a", "W_a", 2 * H, 0, 512, None), ("W_p", "W_p", 5 * TH, 0, 320, None),
]


def _pad4(n):
    return (n + 3) // 4 * 4


class ParamStore:
    def __init__(self, n_items, content_emb, mwdhm, device="cuda"):
        """content_emb [N+1, 250] (row 0 = pad), mwdhm [N, 5] int (month, day, isoweekday, hour+1, minute+1)."""
        self.N = int(n_items)
        self.n_pad = (self.N + 255) // 256 * 256
        self.device = torch.device(device)
        dev = self.device
        content_emb = np.asarray(content_emb, dtype=np.float32)
        if content_emb.shape != (self.N + 1, H):
            raise ValueError(f"content_emb must be [{self.N + 1}, {H}], got {content_emb.shape}")
        mwdhm = np.asarray(mwdhm)
        if mwdhm.shape != (self.N, 5):
            raise ValueError(f"publish_time_MWDHM must be [{self.N}, 5], got {mwdhm.shape}")
        hi = np.array([12, 31, 7, 24, 60])
        if (mwdhm < 0).any() or (mwdhm > hi).any():
            raise ValueError("publish_time_MWDHM out of range for the month/day/week/hour/minute tables")
        self.content = torch.zeros(self.N + 1, HP, device=dev)
        self.content[:, :H] = torch.from_numpy(content_emb).to(dev)
        self.mwdhm = torch.from_numpy(mwdhm.astype(np.int32)).contiguous().to(dev)
        # backing rows padded to a multiple of 8 so that the table splits evenly over 1/2/4/8 data-parallel ranks
        self.rows_alloc = (self.N + 1 + 7) // 8 * 8
        self.item_full = torch.zeros(self.rows_alloc, HP, device=dev)
        self.item_m_full = torch.zeros_like(self.item_full)
        self.item_v_full = torch.zeros_like(self.item_full)
        self.item_g_full = torch.zeros_like(self.item_full)
        self.item, self.item_m = self.item_full[: self.N + 1], self.item_m_full[: self.N + 1]
        self.item_v, self.item_g = self.item_v_full[: self.N + 1], self.item_g_full[: self.N + 1]
        self.iext = torch.zeros(self.n_pad, nv.KEXT, device=dev, dtype=torch.bfloat16)
        # flat small-parameter buffer; every segment starts 16-byte aligned
        offs, off = [], 0
        for _, shp in SMALL:
            offs.append(off)
            off += _pad4(int(np.prod(shp)))
        self.seg_start = offs
        self.seg_len = [int(np.prod(s)) for _, s in SMALL]
        self.flat_size = off
        # clip_by_norm / Adam run per tensor over [start, start+len): pads stay zero because their grads are zero
        seg = []
        for s, l in zip(self.seg_start, self.seg_len):
            seg.append(s)
        seg_off = np.array(self.seg_start + [off], dtype=np.int32)
        self.seg_off = torch.from_numpy(seg_off).to(dev)
        self.theta = torch.zeros(off, device=dev)
        self.theta_m = torch.zeros_like(self.theta)
        self.theta_v = torch.zeros_like(self.theta)
        self.theta_g = torch.zeros_like(self.theta)
        self.sqnorm_small = torch.zeros(len(SMALL), nv.NORM_SPLIT, device=dev)
        self.sqnorm_item = torch.zeros(1, device=dev)
        self.norm_partial = torch.zeros(1184, device=dev)
        self.norm_ticket = torch.zeros(1, device=dev, dtype=torch.int32)
        self.step = torch.zeros(1, device=dev, dtype=torch.int32)
        self.version = 0        # bumped whenever the item table changes (load, train step): cached statistics key on it
        # per-row "already updated in step t" marks of tcar_adam_item_rows (cleared whenever `step` is set from outside)
        self.row_flags = torch.zeros(self.rows_alloc, device=dev, dtype=torch.int32)
        self.w = {n: self._view(self.theta, i) for i, (n, _) in enumerate(SMALL)}
        self.g = {n: self._view(self.theta_g, i) for i, (n, _) in enumerate(SMALL)}
        self.ct_tab = torch.zeros(nv.NBINS, TH, device=dev)
        self.ct_scale = torch.zeros(nv.NBINS, device=dev)
        # pre-split weight copies
        shapes, index = dict(SMALL), {n: i for i, (n, _) in enumerate(SMALL)}
        table, off = [], 0
        self.wp_off, self.wp_pitch = {}, {}
        for name, src, rows, row0, pitch, fold in PREPPED:
            srows, scols = shapes[src]
            self.wp_off[name], self.wp_pitch[name] = off, pitch
            table.append([self.seg_start[index[src]], srows, scols, off + row0 * pitch, pitch,
                          self.seg_start[index[fold]] if fold else -1, srows - shapes[fold][0] if fold else 0])
            off += rows * pitch
        self.wp_table = torch.tensor(table, dtype=torch.int32, device=dev)
        self.wp_hi = torch.zeros(off, device=dev)
        self.wp_lo = torch.zeros(off, device=dev)
        self.wh = {n: self.wp_hi[self.wp_off[n]:] for n, *_ in PREPPED}
        self.wl = {n: self.wp_lo[self.wp_off[n]:] for n, *_ in PREPPED}

    def _view(self, flat, i):
        s, l = self.seg_start[i], self.seg_len[i]
        return flat[s:s + l].view(SMALL[i][1])

    # ------------------------------------------------------------------ initialisation / import / export
    def init_reference(self, emb_stddev=0.002, stddev=0.05, seed=2020):
        """Reference initialisers: embedding tables from the GLOBAL legacy NumPy stream in creation order
        (modules.py:32; main.py:11-12 seeds it with 2020), dense weights ~ N(0, stddev) (modules.py:50-51,65)."""
        p = {}
        for name, sd, zero_pad in [("item", emb_stddev, True), ("pos", 0.02, False), ("month", emb_stddev, True),
                                   ("day", emb_stddev, True), ("week", emb_stddev, True), ("hour", emb_stddev, True),
                                   ("minute", emb_stddev, True), ("dur", emb_stddev, False)]:
            shp = (self.N + 1, H) if name == "item" else dict(SMALL)[name]
            t = np.random.normal(0, sd, shp)
            if zero_pad:
                t[0] = 0.0
            p[name] = t.astype(np.float32)
        g = np.random.default_rng(seed)
        for name, shp in SMALL[7:]:
            p[name] = g.normal(0, stddev, shp).astype(np.float32)
        self.load(p)

    def load(self, params):
        """params: dict name -> array/tensor in the reference shapes (item [N+1,250])."""
        dev = self.device
        item = torch.as_tensor(np.asarray(params["item"], dtype=np.float32))
        self.item.zero_()
        self.item[:, :H] = item.to(dev)
        for n, shp in SMALL:
            self.w[n].copy_(torch.as_tensor(np.asarray(params[n], dtype=np.float32)).view(shp).to(dev))
        for t in (self.item_m, self.item_v, self.theta_m, self.theta_v):
            t.zero_()
        self.step.zero_()
        self.row_flags.zero_()
        self.version += 1
        self.rebuild_iext()
        self.prep_weights()

    def export(self):
        out = {"item": self.item[:, :H].detach().cpu().clone()}
        for n, _ in SMALL:
            out[n] = self.w[n].detach().cpu().clone()
        return out

    def export_grads(self):
        out = {"item": self.item_g[:, :H].detach().cpu().clone()}
        for n, _ in SMALL:
            out[n] = self.g[n].detach().cpu().clone()
        return out

    def prep_weights(self):
        """Refresh the tf32 hi/lo copies of the dense weights (after load and after every Adam step)."""
        nv.counted_call("tcar_prep_weights", 1, nv.ptr(self.theta), nv.ptr(self.wp_table), len(PREPPED),
                        nv.ptr(self.wp_hi), nv.ptr(self.wp_lo))

    def rebuild_iext(self):
        nv.call("tcar_build_iext", nv.ptr(self.item), nv.ptr(self.content), nv.ptr(self.mwdhm), nv.ptr(self.iext),
                self.N, self.n_pad)

    def state_dict(self):
        return {"params": self.export(), "item_m": self.item_m[:, :H].cpu(), "item_v": self.item_v[:, :H].cpu(),
                "theta_m": self.theta_m.cpu(), "theta_v": self.theta_v.cpu(), "step": int(self.step.item())}

    def load_state_dict(self, sd):
        self.load(sd["params"])
        self.item_m[:, :H] = sd["item_m"].to(self.device)
        self.item_v[:, :H] = sd["item_v"].to(self.device)
        self.theta_m.copy_(sd["theta_m"].to(self.device))
        self.theta_v.copy_(sd["theta_v"].to(self.device))
        self.step.fill_(int(sd["step"]))
        self.row_flags.zero_()
