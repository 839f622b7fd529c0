"""Build libfecb200.so in-tree (sm_100a only).

    python finiteelementcontainers.jl_b200/build.py [--force] [--jobs N]

One nvcc invocation per translation unit (run in parallel), then one link step.  The shared
library lands in finiteelementcontainers.jl_b200/lib/ so it travels with the source snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfecb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
METIS = "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a"  # 64-bit idx_t build shipped with the toolkit
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp,-O3", "--expt-relaxed-constexpr"]
# sweeps: FECB200_DEFINES="-DFEC_TE=128 -DFEC_MINB3=3" FECB200_VARIANT=te128_b3 python build.py
DEFINES = os.environ.get("FECB200_DEFINES", "").split()
VARIANT = os.environ.get("FECB200_VARIANT", "")
if VARIANT:
    OBJ = os.path.join(HERE, "build", VARIANT)
    LIB = os.path.join(LIBDIR, f"libfecb200_{VARIANT}.so")

SOURCES = ["api.cu", "aux.cu", "plan.cu", "plan_gpu.cu", "loads.cu", "vmm.cu", "comm.cu", "dispatch_hex8.cu", "dispatch_quad_tri.cu", "dispatch_tet.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "kernel_mat2.cuh", "kernel_mat_scalar.cuh", "physics.cuh", os.path.join("..", "..", "include", "fecb200.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [NVCC, *ARCH, *FLAGS, *DEFINES, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, jobs=None, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    todo = [s for s in SOURCES
            if force or _newer(os.path.join(OBJ, s.replace(".cu", ".o")), [os.path.join(CSRC, s), *hdrs])]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as ex:
            for src, rc, out in ex.map(_compile, todo):
                if verbose:
                    print(f"[fecb200 build] nvcc {src}: {'ok' if rc == 0 else 'FAILED'}", flush=True)
                if rc != 0:
                    sys.stderr.write(out)
                    raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, METIS, "-Xcompiler", "-fopenmp", "-lgomp", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        if verbose:
            print(f"[fecb200 build] linked {LIB}", flush=True)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    a = ap.parse_args()
    build(a.force, a.jobs)
