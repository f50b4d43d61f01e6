#!/usr/bin/env python
"""collect_profiles.py RUN_DIR TAG: copy the measurement set of tools/gpu_final.sh from gpurun_out/RUN_DIR
into profiles/TAG/ (bench lines, ncu raw pages, launch lists; source pages trimmed to the hot lines),
write profiles/TAG/README.md and profiles/traffic.json (DRAM bytes per step of the dominant kernels)."""
import csv, json, os, shutil, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", sys.argv[1])
tag = sys.argv[2]
dst = os.path.join(ROOT, "profiles", tag)
os.makedirs(dst, exist_ok=True)

def load(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[h], rows[h + 1], rows[h + 2:]

def dram_per_kernel(path, match):
    names, units, rows = load(path)
    col = {k: i for i, k in enumerate(names)}
    tot = 0.0; t = 0.0; n = 0
    for r in rows:
        if len(r) < len(names) or match not in r[col["Kernel Name"]]:
            continue
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(r[col[key]].replace(",", ""))
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[col[key]]]
            tot += v
        v = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        t += v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[units[col["gpu__time_duration.sum"]]]
        n += 1
    return tot, t, n

md = [f"# profiles/{tag}: measurement set of `tools/gpu_final.sh` (one B200, this image)\n",
      "Bench lines are the unmodified JSON lines of `bench.py`; `*_raw.csv` is `ncu --set full "
      "--clock-control none` exported with `--page raw --csv`, `*_launches.csv` the "
      "`--metrics gpu__time_duration.sum` launch list of one step of the same command (cold-cache, "
      "serialised: shares, not absolutes), `*_hot_lines.txt` the source page reduced to its hottest lines.\n"]
traffic = {}
for f in sorted(os.listdir(src)):
    p = os.path.join(src, f)
    if f.endswith("_source.csv"):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source.py"), p, "40"],
                             capture_output=True, text=True).stdout
        open(os.path.join(dst, f.replace("_source.csv", "_hot_lines.txt")), "w").write(out)
    elif f.endswith((".json", ".csv", ".log")) and os.path.getsize(p) < 4_000_000:
        shutil.copy(p, os.path.join(dst, f))

md.append("## bench lines\n")
md.append("| file | step ms | GFLOP/s | semiring-kernel ms | algorithmic GB/s (frac of 6551.7) | e2e ms (upload / multiply / fetch) | CPU reference GFLOP/s (threads) |")
md.append("|---|---|---|---|---|---|---|")
for f in sorted(os.listdir(src)):
    if not (f.startswith("bench_") and f.endswith(".json")):
        continue
    try:
        d = json.loads(open(os.path.join(src, f)).read().strip().splitlines()[-1])
    except Exception:
        continue
    if d.get("impl") == "reference":
        md.append(f"| `{f}` | {d['ms_per_step']:.0f} (sample) | {d['value']:.3f} | - | - | - | {d['value']:.3f} ({d['cpu_baseline']['cores']}) |")
        continue
    r = d["roofline"]; e = d.get("e2e") or {}; b = e.get("breakdown") or {}; c = d.get("cpu_baseline") or {}
    e2e = f"{e.get('ms_per_step', 0):.1f} ({b.get('upload_ms', 0):.1f} / {b.get('multiply_ms', 0):.1f} / {b.get('fetch_ms', 0):.1f})" if e else "-"
    cpu = f"{c['value']:.3f} ({c['cores']})" if c.get("value") else "-"
    md.append(f"| `{f}` | {d['ms_per_step']:.3f} | {d['value']:.1f} | {r['kernel_ms']:.3f} | {r['achieved']:.0f} ({r['frac']:.4f}) | {e2e} | {cpu} |")

md.append("\n## ncu captures\n")
for name, match, key in (("tri_s22", "dotg_kernel", "tri_s22"), ("sssp_s22", "spmv_stream", "sssp_s22"),
                         ("bfs_s22", "saxpyv", "bfs_s22"), ("spgemm_rmat16", "saxpy_", "spgemm_rmat_s16")):
    raw = os.path.join(src, f"{name}_raw.csv")
    if not os.path.exists(raw):
        continue
    launches = os.path.join(src, f"{name}_launches.csv")
    args = [sys.executable, os.path.join(ROOT, "tools", "profile_md.py"), name, raw]
    if os.path.exists(launches):
        args.append(launches)
    md.append(subprocess.run(args, capture_output=True, text=True).stdout)
    tot, t, n = dram_per_kernel(raw, match)
    steps = 1                                      # every capture spans the launches of one step
    traffic[key] = int(tot / steps)
    md.append(f"DRAM traffic of `{match}*` ({n} launches captured): {tot / steps / 1e9:.2f} GB per step, "
              f"{t / steps:.2f} ms under ncu.\n")
open(os.path.join(dst, "README.md"), "w").write("\n".join(md) + "\n")
tj = os.path.join(ROOT, "profiles", "traffic.json")
old = json.load(open(tj)) if os.path.exists(tj) else {}
old.update(traffic)
for k in traffic:                                  # a fresh capture supersedes a "not captured" note
    old.pop(k + "_note", None)
json.dump(old, open(tj, "w"), indent=1)
print("\n".join(md)[:3000])
