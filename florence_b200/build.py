"""Builds libflorence_b200.so (hand-written CUDA for sm_100a + the C ABI of include/florence_b200.h) in-tree with nvcc.

    python -m florence_b200.build [--force] [--verbose]

Objects are cached by source mtime under florence_b200/csrc/_obj; the translation units compile in parallel.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libflorence_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("FL_EXTRA_NVCC", "").split()

# (object name, source, extra defines)
UNITS = [("fl_api", "fl_api.cu", []), ("fl_explicit", "fl_explicit.cu", []), ("fl_pattern", "fl_pattern.cu", []),
         ("fl_dirichlet", "fl_dirichlet.cu", []), ("fl_gather", "fl_gather.cu", []), ("fl_stream", "fl_stream.cu", [])] + \
        [("fl_implicit_%d" % k, "fl_implicit.cu", ["-DFL_IMPL_PART=%d" % k]) for k in range(7)]
HEADERS = ["fl_math.cuh", "fl_internal.cuh", "fl_implicit.cuh", "fl_explicit_mma.cuh", "fl_implicit_mma.cuh", "fl_implicit_warp.cuh", "fl_gather.cuh", os.path.join("..", "..", "include", "florence_b200.h")]


def _newest_header():
    return max(os.path.getmtime(os.path.join(CSRC, hname)) for hname in HEADERS if os.path.exists(os.path.join(CSRC, hname)))


def _compile(unit, force, verbose):
    name, src, defs = unit
    srcp = os.path.join(CSRC, src)
    if not os.path.exists(srcp):
        return None
    objp = os.path.join(OBJ, name + ".o")
    if not force and os.path.exists(objp) and os.path.getmtime(objp) >= max(os.path.getmtime(srcp), _newest_header()):
        return objp
    cmd = [NVCC] + FLAGS + defs + ["-c", srcp, "-o", objp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(OBJ, name + ".log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (name, res.stderr[-4000:]))
    if verbose:
        print("compiled", name)
    return objp


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        objs = [o for o in ex.map(lambda u: _compile(u, force, verbose), UNITS) if o]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stderr[-4000:])
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
