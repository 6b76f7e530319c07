"""Per-instruction stall summary from `ncu -i rep --page source --csv --kernel-name regex:<k>` (first launch only)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        if data:
            break
        continue
    data.append(r)
si, sm, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[sm]) for r in data)
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
print("instructions", len(data), "samples", tot)
print("stall totals:", [(k, round(100 * v / tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
print("warp instructions executed:", sum(int(r[ie]) for r in data))
for r in sorted(data, key=lambda r: -int(r[sm]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(r[sm].rjust(6), r[ie].rjust(9), r[si][:64].ljust(64), {hdr[i][6:]: r[i] for i in stalls if int(r[i]) > 0.25 * max(1, int(r[sm]))})
ops = collections.Counter()
for r in data:
    t = r[si].split()
    ops[(t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")).split(".")[0]] += int(r[ie])
print(ops.most_common(14))
