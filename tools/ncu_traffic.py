"""DRAM bytes per launch of every kernel in an `ncu --page raw --csv` dump: dram__bytes_read.sum + dram__bytes_write.sum of the
largest launch of each kernel (a kernel launched on several operands is reported for the biggest one, which is the launch
bench.py brackets).
usage: python tools/ncu_traffic.py raw.csv [raw2.csv ...]  -> JSON {kernel: bytes}; bench.py reads profiles/r02_traffic.json (assembled from the per-workload dumps)"""
import csv, json, re, sys
from collections import defaultdict

def short(name):
    name = re.sub(r"\(.*", "", name); name = re.sub(r"^void ", "", name); name = re.sub(r"<unnamed>::", "", name)
    return re.sub(r"\(anonymous namespace\)::", "", name)

def scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)

out = {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(list)
    for r in rows[2:]:
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[ix[m]].replace(",", "")) * scale(units[ix[m]])
        agg[short(r[ix["Kernel Name"]])].append(b)
    out[path] = {k: max(v) for k, v in agg.items()}
print(json.dumps(out, indent=1))
