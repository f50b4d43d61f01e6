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
