"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name":
            break          # next kernel instance
        continue
    data.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
samples = 0
for r in data:
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
    samples += int(r[ix["# Samples"]] or 0)
print("instructions", len(data), "total samples", samples)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-24s %7d %5.1f%%" % (s, v, 100.0 * v / max(samples, 1)))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
    st = {s: int(r[ix[s]] or 0) for s in stalls}
    main = max(st, key=st.get)
    print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(8), main.ljust(18), r[ix["Source"]].strip()[:100])
