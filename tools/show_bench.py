"""Compact view of a bench.py JSON line: python tools/show_bench.py file.json"""
import json, sys
r = json.load(open(sys.argv[1]))
def rf(o): return f"{o['achieved']:.0f} GB/s = {o['frac']:.3f} of measured ({o['frac_of_8TBs']:.3f} of 8 TB/s), {o['avg_launch_ms']:.4f} ms"
if r.get("value"): print("C3 value", round(r["value"]), "GB/s  ms/step", round(r["ms_per_step"], 2), "roofline", rf(r["roofline"]), r["clocks"])
if r.get("e2e") and r["e2e"].get("value"): print("e2e", round(r["e2e"]["value"], 1), "pcie_bound_frac", round(r["e2e"].get("pcie_bound_frac", 0), 2), "| codes only", round(r["e2e_codes_only"]["value"], 1))
for leg in ("per_token", "c4", "paged"):
    if r.get(leg):
        for k, v in r[leg].items():
            if isinstance(v, dict) and "achieved" in v: print(leg, k, rf(v))
if r.get("decode"):
    for k, v in r["decode"].items():
        if isinstance(v, dict) and "tok_s" in v: print("decode", k, round(v["tok_s"], 1), "tok/s", round(v["ms_per_step"], 2), "ms", rf(v["roofline"]))
        if k == "append_sharded": print("decode append (head-sharded)", v)
for leg in ("c2", "c1"):
    if r.get(leg):
        for k, v in r[leg].items():
            if isinstance(v, dict) and "tok_s" in v: print(leg, k, round(v["tok_s"], 1), "tok/s", v.get("est_kv_cache_mb"))
for k in ("cpu_baseline", "cpu_baseline_reference_torch", "gpu_baseline_reference_torch"):
    if r.get(k): print(k, round(r[k]["value"], 2), r[k].get("cores"))
print("gpu_launches", r.get("gpu_launches"), "n_gpus", r.get("n_gpus"))
