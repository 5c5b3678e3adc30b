"""Turn `ncu --set full` reports into the committed summaries: profiles/<tag>_<name>.md (key metrics per launch)
and profiles/ncu_traffic.json (DRAM bytes per launch, keyed by workload and the kernel label bench.py uses).

    python tools/summarize_ncu.py <tag> <workload> <kernel label> <report.ncu-rep> [...more label report pairs]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

# --out DIR: write there instead of profiles/ (on the GPU box only gpurun_out/ travels back)
OUT = os.path.join(ROOT, "profiles")
if sys.argv[1] == "--out":
    OUT = sys.argv[2]
    del sys.argv[1:3]
tag, workload = sys.argv[1], sys.argv[2]
pairs = list(zip(sys.argv[3::2], sys.argv[4::2]))
traffic_path = os.path.join(OUT, "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
out = ["# ncu --set full --clock-control none, %s, workload %s\n" % (tag, workload)]
for label, rep in pairs:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out.append("## %s  (`%s`)\n" % (label, os.path.basename(rep)))
    tot = []          # (duration, bytes) per captured launch; the LONGEST launch is the one bench.py reports on
    for r in rows[2:]:
        d = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i], units[i])
        out.append("\n".join("- `%s`: %s %s" % (k, v[0], v[1]) for k, v in d.items()) + "\n")
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            tot.append((float(d["gpu__time_duration.sum"][0]), float(rd[0]) * UNIT.get(rd[1], 1) + float(wr[0]) * UNIT.get(wr[1], 1)))
    if tot:
        best = max(tot)
        traffic.setdefault(workload, {})[label] = best[1]
        out.append("DRAM traffic (read + write) of the longest of the %d captured launches: %.1f MB\n" % (len(tot), best[1] / 1e6))
open(os.path.join(OUT, "%s_ncu_full_%s.md" % (tag, workload)), "w").write("\n".join(out))
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
print("\n".join(out)[:2500])
