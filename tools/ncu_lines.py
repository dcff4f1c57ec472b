"""Per-source-line stall samples of one kernel from an .ncu-rep captured with --import-source on (code built with -lineinfo).
usage: python tools/ncu_lines.py rep.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur_file = None
agg = defaultdict(lambda: defaultdict(float)); text = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] == "": continue          # SASS rows carry no line number; the CUDA row above aggregates them
    d = dict(zip(hdr, r))
    key = (cur_file, int(r[0]))
    text[key] = r[1].strip()[:110]
    try: agg[key]["samples"] += float(d.get("# Samples") or 0)
    except ValueError: pass
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: agg[key][k] += float(v or 0)
            except ValueError: pass
tot = sum(a["samples"] for a in agg.values()) or 1.0
print(f"# {rep} {kern}: {tot:.0f} samples")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((v, k[6:]) for k, v in a.items() if k.startswith("stall_") and v > 0), reverse=True)[:3]
    print(f"{100*a['samples']/tot:5.1f}%  {key[0]}:{key[1]:<4d} {' '.join(f'{n}={v:.0f}' for v, n in st):40s} | {text[key]}")
