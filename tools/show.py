#!/usr/bin/env python
"""print the interesting fields of bench.py JSON lines: python tools/show.py file..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r = d["roofline"]
    print(f"{f}\n  {d['config']['workload'][:100]}")
    print(f"  value {d['value']:.2f} GF/s  step {d['ms_per_step']:.3f} ms  device {d['device_ms_per_step']:.3f} ms  "
          f"kernel {r['kernel_ms']:.3f} ms  launches {d['gpu_launches']}  madds {d['config']['madds_per_step']}  nnzT {d['config']['nnz_T']}")
    print(f"  e2e {d['e2e']['value']:.3f} GF/s {d['e2e']['ms_per_step']:.1f} ms   roofline achieved {r['achieved']:.1f} GB/s frac {r['frac']:.4f} "
          f"(whole step {r['step_algo_gbs']:.1f} GB/s)  algoB {r['algorithmic_bytes']}")
    if d.get("cpu_baseline"):
        print("  cpu", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("seconds"))
