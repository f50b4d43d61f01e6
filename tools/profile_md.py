#!/usr/bin/env python
"""profile_md.py TITLE raw.csv [launches.csv] -> markdown summary of an ncu capture (stdout).
raw.csv: `ncu -i x.ncu-rep --page raw --csv`; launches.csv: `ncu --metrics gpu__time_duration.sum --csv`."""
import csv, sys, collections

def load(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[h], rows[h + 1] if len(rows) > h + 1 else [], rows[h + 1:]

def short(name):
    name = name.replace("gb200::", "").replace("void ", "")
    return name[:name.index("(")] if "(" in name else name

title = sys.argv[1]
names, units, rows = load(sys.argv[2])
col = {k: i for i, k in enumerate(names)}
M = [("time ms", "gpu__time_duration.sum", 1), ("DRAM rd GB", "dram__bytes_read.sum", 1), ("DRAM wr GB", "dram__bytes_write.sum", 1),
     ("DRAM %pk", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("SM %pk", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
     ("L2 hit %", "lts__t_sector_hit_rate.pct", 1), ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
     ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1), ("warp-inst G", "smsp__inst_executed.sum", 1e-9),
     ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("regs", "launch__registers_per_thread", 1),
     ("grid", "launch__grid_size", 1), ("block", "launch__block_size", 1)]
def conv(v, unit, key):
    v = float(v.replace(",", ""))
    if key.startswith("gpu__time"):
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)
    if key.startswith("dram__bytes"):
        v *= {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1, "Tbyte": 1e3}.get(unit, 1)
    return v
print(f"### {title}\n")
print("| kernel | " + " | ".join(m[0] for m in M) + " |")
print("|---|" + "---|" * len(M))
for r in rows[1:]:
    if len(r) < len(names):
        continue
    cells = []
    for label, key, scale in M:
        if key in col and r[col[key]] not in ("", "n/a"):
            v = conv(r[col[key]], units[col[key]], key) * scale
            cells.append(f"{v:.3g}" if abs(v) < 1000 else f"{v:.0f}")
        else:
            cells.append("-")
    print(f"| `{short(r[col['Kernel Name']])}` | " + " | ".join(cells) + " |")
if len(sys.argv) > 3:
    names, units, rows = load(sys.argv[3])
    col = {k: i for i, k in enumerate(names)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) < len(names) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(r[col["Metric Unit"]], 1)
        k = short(r[col["Kernel Name"]])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values()) or 1
    print(f"\nLaunch list of one step ({sum(v[0] for v in agg.values())} launches, {tot:.2f} ms under ncu; cold-cache and serialised: shares, not absolutes)\n")
    print("| kernel | launches | ms | share |")
    print("|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |")
