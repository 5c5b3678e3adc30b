"""Group an `ncu --page source --csv` export by SASS opcode: samples, executions, dominant stall reasons."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name": break
        continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("MUFU", "LDG", "STG", "LDS", "STS", "LDTM", "SYNCS", "BAR", "SHFL", "LDL", "STL", "ST", "LD")) and "." in op else "")
    n = int(r[ix["# Samples"]] or 0); e = int(r[ix["Instructions Executed"]] or 0)
    a = agg[op]; a[0] += n; a[1] += e; tot += n
    for s in stalls:
        a[2][s] += int(r[ix[s]] or 0)
print("total samples", tot)
for op, (n, e, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    top = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in c.most_common(3))
    print("%-14s samples %6d (%4.1f%%) exec %9d   %s" % (op, n, 100.0 * n / max(tot, 1), e, top))
