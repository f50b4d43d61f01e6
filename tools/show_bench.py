#!/usr/bin/env python
"""show_bench.py file.json...: one line per bench.py result"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        try:
            err = open(f.replace(".json", ".err")).read()[-600:]
        except Exception:
            err = ""
        print(f"{f}: ERR {e} {err}")
        continue
    e2e = d.get("e2e") or {}
    r = d["roofline"]
    print(f"{f.split('/')[-1]:28s} step {d['ms_per_step']:8.3f} ms  {d['value']:8.1f} GF/s  kernel {r['kernel_ms']:8.3f} ms  frac {r['frac']:.4f}  "
          f"e2e {e2e.get('ms_per_step', 0):8.2f} ms {e2e.get('breakdown')}  launches {d.get('gpu_launches')}")
    for nb in d.get("neighbours") or []:
        if "error" in nb:
            print(f"    neighbour {nb.get('metric')}: ERROR {nb['error'][-200:]!r}")
        elif "cases" in nb:
            cs = ", ".join(f"{k} {v['best_ms']:.3f} ms (frac {v['roofline']['frac']:.3f})" for k, v in nb["cases"].items())
            cpu = nb.get("cpu_baseline") or {}
            ref = ", ".join(f"{k} {cpu[k]['value']:.0f} ms" for k in ("same", "union") if k in cpu)
            print(f"    neighbour {nb['metric']}: {cs}; parity {nb.get('parity', {}).get('identical')}; reference on the host ({cpu.get('sample', '-')}): {ref}")
        else:
            cpu = nb.get("cpu_baseline") or {}
            print(f"    neighbour {nb['metric']}: {nb.get('best_ms', 0):.3f} ms for {nb.get('nnz')} entries, frac {nb.get('roofline', {}).get('frac', 0):.3f}, "
                  f"parity {nb.get('parity', {}).get('identical')}; reference on the host: {cpu.get('value', 0):.0f} ms ({cpu.get('sample', '-')})")
