#!/usr/bin/env python
"""per-kernel mean duration of an ncu launch list (gpurun_out/launches_*.csv)"""
import collections, csv, io, sys
for f in sys.argv[1:]:
    lines = [l for l in open(f) if l.startswith('"')]
    per = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        n = r["Kernel Name"].replace("void ", "").replace("nbm::", "").replace("stencil_tma::", "").split("(")[0].split("<")[0]
        per.setdefault(n, []).append(float(r["Metric Value"]) / 1e3)
    print(f)
    for k, v in per.items():
        print(f"  {k:28s} n={len(v):3d} mean {sum(v)/len(v):8.1f} us  min {min(v):8.1f}")
