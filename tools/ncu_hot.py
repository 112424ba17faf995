"""Summarise an `ncu --page source --csv` export: top SASS instructions by stall samples with their dominant stall
reasons. Usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    data.append((n, r))
total = sum(n for n, _ in data) or 1
print(f"total samples {total}; instructions {len(data)}")
agg = {s: 0 for s in stalls}
for n, r in data:
    for s in stalls:
        try:
            agg[s] += int(r[col[s]] or 0)
        except ValueError:
            pass
print("stall mix:", ", ".join(f"{s[6:]} {100 * v / total:.1f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for idx, (n, r) in enumerate(data):
    r.append(idx)
for n, r in sorted(data, key=lambda t: -t[0])[:top]:
    rs = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{100 * n / total:5.1f}%  #{r[-1]:5d} {r[col['Source']][:70]:70s} ex={r[col['Instructions Executed']]:>9s} " +
          " ".join(f"{s}:{v}" for v, s in rs if v))
