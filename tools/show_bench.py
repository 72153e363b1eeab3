import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            print("==", f, "gpus", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches/step", d["launches_per_step"], "lookahead", d["config"].get("lookahead"))
            print("   t20", round(d["t20"]["value"]), "loop", round(d["train_loop"]["value"]) if d.get("train_loop") else None, "devloop", round(d["train_loop"]["device_sampler"]["value"]) if d.get("train_loop") else None)
            print("   parity", d["parity"] and {k: v for k, v in d["parity"].items() if k != "what"})
            print("   spread", d.get("ms_per_step_spread"))
            print("   e2e_host", d["e2e"].get("host_step_wall"), "kernels", {k: round(v["ms"] * 1e3, 1) for k, v in (d.get("kernels") or {}).items()})
            print("   eval", round(d["eval"]["value"]), "ms", round(d["eval"]["ms_per_step"], 4), "uncert", d["eval"]["uncertified_share"], "e2e", round(d["eval"]["e2e"]["value"]), "qp", d["eval_query_parallel"] and round(d["eval_query_parallel"]["value"]))
