"""profiles/<tag>_launch_shares.md from an `ncu --metrics gpu__time_duration.sum --csv` launch list: the kernels of
ONE diffusion step (between two stage_x_kernel launches), their counts, summed durations and shares."""
import collections
import csv
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
names = [r[ix["Kernel Name"]] for r in data]
vals = [float(r[ix["Metric Value"]]) for r in data]
unit = data[0][ix["Metric Unit"]]
st = [i for i, n in enumerate(names) if "stage_x" in n]
segs = [(a, b) for a, b in zip(st[:-1], st[1:])]


def table(a, b, title):
    tot = sum(vals[a:b])
    agg = collections.OrderedDict()
    for n, v in zip(names[a:b], vals[a:b]):
        k = n.split("(")[0]
        agg.setdefault(k, [0.0, 0])
        agg[k][0] += v
        agg[k][1] += 1
    out = ["## %s: %d launches, sum %.1f %s\n" % (title, b - a, tot, unit), "| kernel | launches | time (%s) | share |" % unit, "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append("| `%s` | %d | %.1f | %.1f %% |" % (k, v[1], v[0], 100 * v[0] / tot))
    return out


# a regular (bf16, tensor-core) diffusion step, and -- when the schedule has one -- the ill-conditioned leading step whose
# U-Net runs in the fp32 sibling (once per plan)
regular = next(((a, b) for a, b in segs if any("conv_chain" in n or "conv_t3" in n for n in names[a:b])), segs[0])
sibling = next(((a, b) for a, b in segs if any("conv_tf32" in n or "conv_f32" in n for n in names[a:b])), None)
out = ["# Kernel shares of one diffusion step (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n",
       "Command: `%s` (the un-profiled step takes less: kernels overlap their set-up via PDL and run warm).\n" % cmd]
out += table(regular[0], regular[1], "regular diffusion step (every step but the first of a plan)")
if sibling:
    out += [""] + table(sibling[0], sibling[1], "ill-conditioned leading step: U-Net of the fp32 sibling (TF32 tensor cores) + our step kernel, once per plan")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
