"""Per-source-line summary of an ncu report: python tools/ncu_lines.py report.ncu-rep [kernel-index] [top]
(ncu --page source --csv --print-source cuda,sass; needs -lineinfo).  Prints, per CUDA source line, the executed warp
instructions, their share, stall samples and the dominant stall reasons."""
import csv, io, subprocess, sys
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# sections: a header row starting with "Line No" begins each (file, kernel) table
tables, cur = [], None
for r in rows:
    if r and r[0] == "Line No":
        cur = {"hdr": r, "rows": []}
        tables.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
big = sorted(tables, key=lambda t: -len(t["rows"]))
t = big[kidx]
h = t["hdr"]
ix = {n: i for i, n in enumerate(h)}
inst_i, samp_i = ix["Instructions Executed"], ix["# Samples"]
stall_cols = [(n, i) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n]
agg = {}
for r in t["rows"]:
    try:
        line = int(r[0])
    except ValueError:
        continue
    a = agg.setdefault(line, {"src": r[1], "inst": 0, "samp": 0, "st": {}})
    a["inst"] += int(float(r[inst_i] or 0))
    a["samp"] += int(float(r[samp_i] or 0))
    for n, i in stall_cols:
        v = int(float(r[i] or 0))
        if v:
            a["st"][n] = a["st"].get(n, 0) + v
tot_i = sum(a["inst"] for a in agg.values()) or 1
tot_s = sum(a["samp"] for a in agg.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = ", ".join(f"{n[6:]} {v}" for n, v in sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
    print(f"L{line:4d} inst {100 * a['inst'] / tot_i:5.1f}% samp {100 * a['samp'] / tot_s:5.1f}%  [{st}]  {a['src'].strip()[:100]}")
