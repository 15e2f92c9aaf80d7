"""In-tree build of the C-ABI library `libnbm_b200.so` (sm_100a only) with nvcc.

`python -m jax_dips_b200.build` or `__graft_entry__.build()`.  The .so stays next to the sources
(git-ignored) so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnbm_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]

# (source, extra flags).  The per-level geometry kernels are built without FMA contraction so that
# fp32 sign/ordering decisions are taken on plainly rounded values (see nbm_geometry.cu).
UNITS = [
    ("nbm_geometry.cu", ["-fmad=false"]),
    ("nbm_step.cu", ["-Xptxas", "-v"]),
]


def _newer(src: str, dst: str) -> bool:
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # sources only: the generated objects (and their ptxas logs) must not make every unit look stale
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "nbm_b200.h"))
    logdir = os.path.join(HERE, "..", "build", "ptxas")
    os.makedirs(logdir, exist_ok=True)
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        if force or any(_newer(d, o) for d in deps):
            cmd = [nvcc] + ARCH + [f for f in COMMON if not f.startswith("--use_fast_math")] + extra + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
            # ptxas -v output (registers / spills per kernel): kept outside the tracked tree; copies that are
            # evidence live under profiles/
            with open(os.path.join(logdir, src.replace(".cu", ".ptxas.log")), "w") as f:
                f.write(r.stdout + r.stderr)
        objs.append(o)
    if force or any(_newer(o, LIB) for o in objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
