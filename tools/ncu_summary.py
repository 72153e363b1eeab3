"""Summarise an `ncu --set full` report of `bench.py --profile_region` into profiles/: per-launch time, DRAM traffic,
DRAM / tensor-pipe utilisation, registers -> <out>.json + <out>.md, and the per-kernel DRAM traffic table
profiles/traffic.json that bench.py copies into `roofline.traffic` / `kernels[*].traffic`.
usage: python tools/ncu_summary.py gpurun_out/prof_step_raw.csv [gpurun_out/prof_bwd_q_raw.csv] --out profiles/r2_a_ncu_full
(.ncu-rep files are accepted too)"""
import argparse, csv, io, json, os, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("reports", nargs="+")
ap.add_argument("--out", required=True)
a = ap.parse_args()

COLS = {"time_us": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread"}
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows_out = []
for rep in a.reports:
    if rep.endswith(".csv"):           # raw page exported on the GPU box (a full-set .ncu-rep exceeds gpurun's 64 MiB)
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("tcar::", ""),
             "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for k, m in COLS.items():
            if m not in idx or r[idx[m]] in ("", "n/a", "no data"):
                d[k] = None
                continue
            v = float(r[idx[m]].replace(",", ""))
            u = units[idx[m]]
            d[k] = v * SCALE.get(u, 1.0)
        rows_out.append(d)
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
with open(a.out + ".json", "w") as f:
    json.dump(rows_out, f, indent=1)
with open(a.out + ".md", "w") as f:
    f.write("| kernel | grid | block | time (us) | DRAM read (MB) | DRAM write (MB) | DRAM % | tensor pipe % | regs |\n|---|---|---|---|---|---|---|---|---|\n")
    for d in rows_out:
        fm = lambda v, s=1.0, p=1: "n/a" if v is None else f"{v / s:.{p}f}"
        f.write(f"| {d['kernel']} | {d['grid']} | {d['block']} | {fm(d['time_us'])} | {fm(d['dram_read'], 1e6)} | "
                f"{fm(d['dram_write'], 1e6)} | {fm(d['dram_pct'])} | {fm(d['tensor_pct'])} | {fm(d['regs'], 1.0, 0)} |\n")
# bench.py kernel name -> ncu kernel name (first launch of the train step, eval-mode forward for score_fwd_eval)
MAP = {"score_fwd_train": "score_fwd_pair_kernel<0>", "score_fwd_eval": "score_fwd_pair_kernel<1>",
       "score_bwd_q": "score_bwd_q_kernel", "score_bwd_i": "score_bwd_i_kernel", "adam_item": "adam_item_kernel",
       "gather_fwd": "gather_fwd_kernel", "pool_fwd": "pool_fwd_kernel", "scatter_add_rows": None,
       "eval_topk": "eval_topk_kernel", "sqnorm_item_grad": None}
traffic = {}
for k, name in MAP.items():
    if name is None:
        continue
    names = {"score_bwd_i_kernel": ("score_bwd_i_tma_kernel<0>", "score_bwd_i_tma_kernel<false>", "score_bwd_i_kernel")}.get(name, (name,))
    hit = [d for d in rows_out if d["kernel"] in names and d["dram_read"] is not None]
    traffic[k] = (hit[0]["dram_read"] + hit[0]["dram_write"]) if hit else None
sc = [d for d in rows_out if d["kernel"].startswith("scatter_") and d["dram_read"] is not None]
if sc:
    n = max(1, len([d for d in sc if d["kernel"] == "scatter_accum_kernel"]))
    traffic["scatter_add_rows"] = sum(d["dram_read"] + d["dram_write"] for d in sc) / n
traffic["sqnorm_item_grad"] = None
traffic["_source"] = (f"{os.path.basename(a.out)}.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, `ncu --set full "
                      "--clock-control none` capture of bench.py --profile_region (B=512, T=20, N=364047); scatter_add_rows = "
                      "its three kernels; score_bwd_q from the lighter-section capture when the full set cannot launch it")
with open(os.path.join(os.path.dirname(a.out) or ".", "traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print(open(a.out + ".md").read())
