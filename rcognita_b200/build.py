"""In-tree build of ``librcg_b200.so`` (hand-written sm_100a CUDA kernels + the C ABI).

``python -m rcognita_b200.build`` or ``rcognita_b200.build.build_library()``.  nvcc
cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "librcg_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INCLUDE]
# rk45.cu: no FMA contraction -- the step-size arithmetic must round like numpy's
SOURCES = {
    "api.cu": [],
    "rk45.cu": ["-fmad=false"],
    "actor.cu": [],
    "actor_ni_f64.cu": [], "actor_3w_f64.cu": [], "actor_2t_f64.cu": [],
    "actor_ni_f32.cu": [], "actor_3w_f32.cu": [], "actor_2t_f32.cu": [],
    "actor_tab_ni.cu": [], "actor_tab_3w.cu": [],
    "actor_opt.cu": [], "actor_opt_ni.cu": [], "actor_opt_3w.cu": [], "actor_opt_2t.cu": [], "actor_ilqr.cu": [],
    "actor_optq_ni.cu": [], "actor_optq_3w.cu": [],
    "critic.cu": ["-fmad=false"],
    "nominal.cu": ["-fmad=false"],
    "critic_fit.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _includes(path: str, seen=None) -> set:
    """Project headers a source reaches through `#include "..."` (csrc/ and include/), transitively."""
    seen = set() if seen is None else seen
    try:
        with open(path) as fh:
            text = fh.read()
    except OSError:
        return seen
    for line in text.splitlines():
        line = line.strip()
        if line.startswith('#include "'):
            name = line.split('"')[1]
            for base in (CSRC, INCLUDE):
                cand = os.path.join(base, name)
                if os.path.exists(cand) and cand not in seen:
                    seen.add(cand)
                    _includes(cand, seen)
    return seen


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        hdrs = sorted(_includes(s)) + [os.path.abspath(__file__)]        # only the headers this source actually includes
        if force or _stale(o, [s] + hdrs):
            jobs.append((src, [nvcc, *ARCH, *COMMON, *extra, "-c", s, "-o", o]))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r

    with ThreadPoolExecutor(max_workers=max(1, min(os.cpu_count() or 4, len(jobs)))) as ex:
        for name, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- nvcc {name}\n{r.stdout}{r.stderr}\n")
            with open(os.path.join(BUILD, name + ".ptxas.log"), "w") as fh:
                fh.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {name}")
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
