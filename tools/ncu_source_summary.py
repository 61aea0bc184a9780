"""Opcode mix and stall-sample summary of an `ncu --page source --csv` export (one kernel)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Source" in r][0]
hdr = rows[hi]
ci, ei, si = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, st, samp = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= max(ci, ei, si):
        continue
    try:
        n = int(r[ei])
    except ValueError:
        continue
    toks = r[ci].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ops[op] += n
    tot += n
    samp[op] += int(r[si] or 0)
    for i in stalls:
        st[hdr[i]] += int(r[i] or 0)
print("warp instructions", tot)
ts = sum(samp.values())
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print(f"{k:26s}{v:11d} {100 * v / tot:5.1f}%   samples {100 * samp[k] / max(ts, 1):5.1f}%")
tt = sum(st.values())
print(" ".join(f"{k[6:]}={100 * v / tt:.1f}%" for k, v in st.most_common(10)))
