#!/usr/bin/env python
"""Print the stage and per-kernel figures of bench.py JSON lines. usage: tools/bench_brief.py file.json ..."""
import json
import os
import sys

for f in sys.argv[1:]:
    if not os.path.exists(f):
        continue
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        e = d.get("e2e", {})
        print(f"== {f}: {d.get('value', 0):.4g} {d.get('unit')}  {d.get('ms_per_step', 0):.4f} ms/step | e2e {e.get('value', 0):.4g} "
              f"{e.get('ms_per_step', 0):.4f} ms | launches/step {d.get('gpu_launches_per_step')} | clocks {d.get('clocks')}")
        print("   stages:", {k: round(v, 4) for k, v in (d.get("stages_ms") or {}).items() if v})
        for r in d.get("roofline_kernels") or []:
            print(f"   {r['kernel']:<22} {r['ms_per_launch']:9.4f} ms  {r['bound']:<6} {r['achieved']:9.2f} {r['unit']:<8} frac {r['frac']:.3f}")
        print("   selection:", d.get("selection"))
