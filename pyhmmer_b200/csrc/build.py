"""Build libb2h.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The shared library is written next to the package (``pyhmmer_b200/libb2h.so``) so that it
travels with the repository snapshot to the GPU box.  Only nvcc + g++ are needed; there is no
dependency on torch, pybind11 or the reference.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libb2h.so")
OBJ = os.path.join(HERE, "_obj")

SOURCES = [
    "b2h_host.cpp",
    "b2h_device.cu",
    "b2h_api.cu",
    "b2h_msv.cu",
    "b2h_dp.cu",
    "b2h_dpreg.cu",
    "b2h_envelope.cu",
    "b2h_generic.cu",
    "b2h_longtarget.cu",
    "b2h_ltvit.cu",
    "b2h_search.cu",
    "b2h_domaindef.cpp",
    "b2h_pressed.cpp",
]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-I", os.path.join(ROOT, "include"), "-I", HERE,
          "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function"]


# the host-side domain definition is written as unit-stride loops for the compiler's vectoriser; its hot functions carry
# target_clones("avx2", "default") (runtime dispatch: no AVX2 requirement on the host); `omp simd` only licenses the
# re-association of the marked float reductions
EXTRA = {"b2h_domaindef.cpp": ["-Xcompiler", "-fopenmp-simd"],
         "b2h_generic.cu": ["-fmad=false"]}     # log-space sums must stay plain IEEE adds


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the CUDA extension")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(ROOT, "include", "b2h.h")] + \
              [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))]
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers + [os.path.abspath(__file__)]):
            cmd = [nvcc] + ARCH + COMMON + EXTRA.get(src, []) + ["-Xptxas", "-v" if verbose else "-O3", "-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stdout))
        if verbose and r.stdout:
            sys.stderr.write(r.stdout)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(OUT, objs):
        run([nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart_static", "-lrt", "-ldl", "-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
