"""Build liberfnet_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m erfnet_pytorch_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "liberfnet_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


TRACE_LIB_PATH = os.path.join(LIB_DIR, "liberfnet_b200_trace.so")
TRACED = ("conv1d_tc.cu", "conv1d_pair_tc.cu", "conv1d_wgrad_tc.cu", "bn_fused.cu")   # sources with per-CTA timeline stamps (compiled in only for the trace library)


def build(force: bool = False, verbose: bool = False) -> str:
    """liberfnet_b200.so (timeline stamps compiled out: ERF_NO_TIMELINE) and liberfnet_b200_trace.so (the same objects except the
    instrumented kernels, for the tools under ERF_TIMELINE=1)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(HERE), "include", "erfnet_b200.h")]
    # a change of the compile flags (or of this recipe) rebuilds everything
    stamp = os.path.join(build_dir, "flags.txt")
    recipe = " ".join(FLAGS) + " | main: -DERF_NO_TIMELINE | trace: " + " ".join(TRACED)
    if not os.path.exists(stamp) or open(stamp).read() != recipe:
        force = True
    objs, trace_objs, jobs = [], [], []
    for s in srcs:
        o = os.path.join(build_dir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([NVCC] + FLAGS + ["-DERF_NO_TIMELINE", "-c", s, "-o", o])
        if os.path.basename(s) in TRACED:
            ot = o[:-2] + "_trace.o"
            trace_objs.append(ot)
            if force or _stale(ot, [s] + hdrs):
                jobs.append([NVCC] + FLAGS + ["-c", s, "-o", ot])
        else:
            trace_objs.append(o)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed for {cmd[-3]}")
                with open(cmd[-1][:-2] + ".ptxas.log", "w") as f:
                    f.write(r.stderr)
    with open(stamp, "w") as f:
        f.write(recipe)
    for path, members in ((LIB_PATH, objs), (TRACE_LIB_PATH, trace_objs)):
        if jobs or force or _stale(path, members):
            cmd = [NVCC, "-shared", "-o", path] + members + ["-gencode", "arch=compute_100a,code=sm_100a"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
