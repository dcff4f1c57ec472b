#!/bin/bash
# driver-shaped visit: smoke, the default bench line, (optionally) the GPU tests
TAG=${1:-fin}; O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
T0=$(date +%s); timeout 900 python bench.py > $O/bench_default_$TAG.json 2> $O/bench_default_$TAG.err; echo "bench wall $(( $(date +%s) - T0 )) s"; tail -3 $O/bench_default_$TAG.err
python tools/bench_brief.py $O/bench_default_$TAG.json | head -16 | cut -c1-300
python - <<P
import json
d=json.loads(open("$O/bench_default_$TAG.json").read().strip().splitlines()[-1])
print("roofline", {k: d["roofline"][k] for k in ("kernel","bound","achieved","peak","frac")})
print("e2e", d["e2e"]["ms_per_step"], "chain", d["e2e_chain"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"])
for k,x in d["configs"].items(): print(k, round(x["ms_per_step"],3), round(x["e2e"]["ms_per_step"],3), x["roofline"]["kernel"], round(x["roofline"]["frac"],3))
s=d["sharded_weight_update"]; print("C4", s["ms_per_step"], s["roofline"]["frac"], s.get("max_rel_err_vs_oracle"))
P
if [ "$2" = "tests" ]; then timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3; fi
