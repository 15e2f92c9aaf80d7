#!/usr/bin/env python
"""Table of the named-configuration runs (gpurun_out/r2_named/*.json -> markdown on stdout)."""
import glob, json, os, re, sys
d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_named"
rows = []
for f in sorted(glob.glob(os.path.join(d, "*gpu.json"))):
    m = re.match(r"(.*)_(\d)gpu\.json", os.path.basename(f))
    line = None
    for l in open(f):
        if l.startswith("{"):
            line = json.loads(l)
    if line is None:
        rows.append((m.group(1), int(m.group(2)), None))
        continue
    rows.append((m.group(1), int(m.group(2)), line))
print("| workload | GPUs | points/s | ms/step | e2e points/s | step frac of FP32 peak (per GPU) | parity_check | slabs (planes) |")
print("|---|---|---|---|---|---|---|---|")
for name, n, L in sorted(rows, key=lambda r: (r[0], r[1])):
    if L is None:
        print(f"| {name} | {n} | failed | | | | | |")
        continue
    pc = L.get("parity_check")
    pcs = "-" if pc is None else (f"ok (peer vs NCCL {pc['peer_vs_nccl_rel']:.1e}, additivity {pc.get('slab_additivity_rel', float('nan')):.1e}, bitwise across ranks)" if pc["ok"] else "FAILED")
    pr = L.get("per_rank")
    slabs = "-" if not pr else "/".join(str(r.get("planes", "?")) for r in pr)
    roof = L.get("roofline") or {}
    sf = roof.get("step", {}).get("frac")
    print(f"| {name} | {n} | {L['value']:.3e} | {L['ms_per_step']:.4f} | {L['e2e']['value']:.3e} | {sf:.3f} | {pcs} | {slabs} |")
