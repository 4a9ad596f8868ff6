"""Builds the in-tree CUDA library  diff_gaussian_rasterization/libgsr_b200.so  for sm_100a.

    python gs-slam-analytica_jacobian_b200/build.py [--force] [--verbose]

Plain nvcc (no torch headers in any translation unit -> seconds, the reference's torch binding TU
takes > 8 minutes).  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "diff_gaussian_rasterization", "libgsr_b200.so")
SOURCES = ["preprocess.cu", "binning.cu", "render.cu", "render_backward.cu", "preprocess_backward.cu", "slam_ops.cu", "window_reduce.cu", "api.cu"]
HEADERS = ["gsr_common.cuh", "gsr_params.h", "render_common.cuh", "tile_sort.cuh", os.path.join("..", "..", "include", "gsr_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC"]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s and %s is missing/stale" % (NVCC, OUT))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    # several ranks of one job may find the library stale at the same moment (torchrun on a box whose snapshot lost the mtimes):
    # one of them builds, the others wait for the lock and find the fresh library
    import fcntl

    with open(os.path.join(objdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return OUT
            return _build_locked(objdir, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(objdir, verbose):
    extra = ["-Xptxas", "-v"] if verbose else []
    extra += os.environ.get("GSR_EXTRA_NVCC_FLAGS", "").split()      # experiments, e.g. -DGSR_BWD_SLOTS=8
    if os.environ.get("GSR_PHASE_PROBE"):      # diagnostic build (tools/phase_probe.py), never the product
        extra += ["-DGSR_PHASE_PROBE"]

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    tmp = OUT + ".tmp%d" % os.getpid()      # link beside the target, then rename: a process loading the library never sees half of it
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
