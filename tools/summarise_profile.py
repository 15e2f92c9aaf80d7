#!/usr/bin/env python
"""Turn the raw outputs of tools/profile.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarise_profile.py r1b

Reads gpurun_out/launches_<tag>.csv (ncu launch list), gpurun_out/prof_<tag>.ncu-rep (ncu --set full of the four
step kernels) and gpurun_out/bench_<tag>_1gpu.json (the live bench line of the same build).
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1b"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

STEP = ["prep_params_kernel", "fwd_nodes_kernel", "extrap_kernel", "residual", "stencil_tma", "merge_lists", "irregular_fb", "irregular_fwd_kernel", "adjoint",
        "irregular_bwd_kernel", "extrap_bwd_kernel", "node_grad", "precond_kernel", "reduce_partials_kernel",
        "apply_update_kernel", "finalize_step_kernel"]


def short(name):
    n = name.replace("void ", "").replace("nbm::", "").replace("stencil_tma::", "")
    return n.split("(")[0].split("<")[0]


# ---- launch list ---------------------------------------------------------------------------------
lines = [l for l in open(os.path.join(G, f"launches_{tag}.csv")) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
per = collections.OrderedDict()
for r in rows:
    per.setdefault(short(r["Kernel Name"]), []).append(float(r["Metric Value"]) / 1e3)
shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{tag}_launches.csv"))
bench = json.loads(open(os.path.join(G, f"bench_{tag}_1gpu.json")).read().strip().splitlines()[-1])
shutil.copy(os.path.join(G, f"bench_{tag}_1gpu.json"), os.path.join(P, f"{tag}_bench_1gpu.json"))
step = [(k, v) for k, v in per.items() if any(k.startswith(s) for s in STEP)]
tot = sum(sum(v) / len(v) for _, v in step)
out = [f"# {tag} - ncu launch list (sphere 256^3, 1 x B200)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 4 --warmup 3 "
       "--no-cpu-baseline --no-graph` (tools/profile.sh).  Raw list: `%s_launches.csv`.  ncu times are cold-cache and "
       "serialised: compare SHARES with the live CUDA-event numbers of `bench.py`." % tag, "",
       "| kernel | launches | mean us (ncu) | share of step (ncu) |", "|---|---|---|---|"]
for k, v in step:
    m = sum(v) / len(v)
    out.append(f"| `{k}` | {len(v)} | {m:.1f} | {100 * m / tot:.1f} % |")
out.append(f"| sum | | {tot:.1f} | |")
sm = bench["roofline"]["stage_ms"]
ms = bench["ms_per_step"]
out += ["", f"## the same build, live (bench.py, CUDA events on the launching stream, `{tag}_bench_1gpu.json`)", "",
        f"step {ms * 1e3:.0f} us, {bench['value']:.3e} points/s (e2e through the C ABI with host buffers "
        f"{bench['e2e']['value']:.3e}); stages (each timed alone, list kernels included): "
        + ", ".join(f"{k} {v * 1e3:.0f} us ({100 * v / sum(sm.values()):.0f} %)" for k, v in sm.items()) + ".", "",
        f"node_grad: {bench['roofline']['achieved']:.1f} TFLOP/s algorithmic = {100 * bench['roofline']['frac']:.0f} % "
        f"of the FP32 FMA peak measured in the same process ({bench['roofline']['peak']:.1f} TFLOP/s); whole step "
        f"{100 * bench['roofline']['step']['frac']:.0f} %; residual+adjoint stages "
        f"{bench['roofline']['hbm']['achieved']:.0f} GB/s = {100 * bench['roofline']['hbm']['frac']:.0f} % of measured HBM."]
open(os.path.join(P, f"{tag}_launch_list_summary.md"), "w").write("\n".join(out) + "\n")

# ---- ncu --set full ------------------------------------------------------------------------------
rep = os.path.join(G, f"prof_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_ncu_full_raw.csv"), "w").write(raw)
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rr[0], rr[1], rr[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
cols = [c for c in cols if c in ix]
md = [f"# {tag} - `ncu --set full --clock-control none` of the step kernels (sphere 256^3, "
      f"{bench['roofline']['nodes_per_launch']} lattice nodes per launch)", "",
      f"Raw metrics: `{tag}_ncu_full_raw.csv`; per-instruction stall samples of node_grad: `{tag}_ncu_node_grad_source.csv`.",
      "", "| kernel | " + " | ".join(c.replace("__", ".").replace(".avg.pct_of_peak_sustained_active", " %").replace(
          ".avg.pct_of_peak_sustained_elapsed", " %").replace("smsp.average_warps_issue_stalled_", "stall ").replace(
          "_per_issue_active.ratio", "") for c in cols) + " |", "|---|" + "---|" * len(cols)]
traffic = None
for r in data:
    name = short(r[ix["Kernel Name"]])
    md.append(f"| `{name}` | " + " | ".join(r[ix[c]] for c in cols) + " |")
    if name.startswith("node_grad"):
        def val(c):
            v, u = float(r[ix[c]]), units[ix[c]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        traffic = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
md += ["", "Units: " + ", ".join(f"{c.split('.')[0].split('__')[-1]} [{units[ix[c]]}]" for c in cols[:3]) +
       "; the rest % of peak (sustained) / counts / stall cycles per issued instruction.", ""]
if traffic:
    md.append(f"node_grad DRAM traffic per launch (read + write): {traffic / 1e6:.1f} MB "
              f"(`roofline.traffic` of bench.py).")
open(os.path.join(P, f"{tag}_ncu_summary.md"), "w").write("\n".join(md) + "\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:node_grad"], capture_output=True,
                     text=True).stdout
open(os.path.join(P, f"{tag}_ncu_node_grad_source.csv"), "w").write(src)
print("\n".join(md[-12:]))
print("traffic bytes:", traffic)
