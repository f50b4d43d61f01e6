#!/usr/bin/env python
"""Hot lines of an `ncu --page source --csv --print-source cuda,sass` export: aggregates the SASS
rows by (file, CUDA line) and prints the top lines by instructions executed and by stall samples."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
# sections: "File Path",... / "Function Name",... / header / rows
cur_file = None; hdr = None; func = None
agg = {}
srcs = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": func = r[1][:60]; continue
    if r[0] == "Line No": hdr = {k: i for i, k in enumerate(r)}; hdr_names = r; continue
    if hdr is None or len(r) < 10: continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (func, cur_file, ln)
    a = agg.setdefault(key, collections.Counter())
    srcs[key] = r[1].strip()[:100]
    for name in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_branch_resolving", "L1 Wavefronts Shared", "L2 Theoretical Sectors Global", "Divergent Branches"):
        # the CUDA-view row of a line already aggregates its SASS rows; take the row's own numbers
        i = hdr.get(name)
        if i is not None and i < len(r) and r[i] not in ("", "-"):
            try: a[name] = max(a[name], float(r[i]))
            except ValueError: pass
tot_i = sum(a["Instructions Executed"] for a in agg.values()) or 1
tot_s = sum(a["# Samples"] for a in agg.values()) or 1
print(f"total inst {tot_i:.3g}  samples {tot_s:.0f}")
for title, k in (("by instructions executed", "Instructions Executed"), ("by stall samples", "# Samples")):
    print("--", title)
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][k])[:top]:
        ti = a["Thread Instructions Executed"] / max(a["Instructions Executed"], 1)
        print(f"{key[1]:>16s}:{key[2]:<4d} inst {100*a['Instructions Executed']/tot_i:5.1f}% thr/inst {ti:4.1f} smp {100*a['# Samples']/tot_s:5.1f}% "
              f"longsb {a['stall_long_sb']:.0f} bar {a['stall_barrier']:.0f} shortsb {a['stall_short_sb']:.0f} wait {a['stall_wait']:.0f} br {a['stall_branch_resolving']:.0f} | {srcs[key]}")
