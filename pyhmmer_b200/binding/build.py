"""Build pyhmmer_cuda (the Cython binding of libb2h.so into an installed pyhmmer) in-tree.

    python pyhmmer_b200/binding/build.py [path of the directory that holds the installed `pyhmmer` package]

What a pyhmmer maintainer's CMake would do with find_package(PyHMMER) (src/cmake/PyHMMERConfig.cmake.in): Cython with
PyHMMER_CYTHON_DIRS on its include path and the reference's compile-time constant HMMER_IMPL, the C compiler with
PyHMMER_INCLUDE_DIRS, the linker with PyHMMER_LIBRARIES -- plus include/b2h.h and libb2h.so.  The default pyhmmer is the
unmodified reference installed under baseline/_ref; the module is written next to this file.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)


def build(site=None, force=False):
    site = os.path.abspath(site or os.path.join(ROOT, "baseline", "_ref"))
    libs = os.path.join(site, "pyhmmer.libs")
    if not os.path.isdir(os.path.join(site, "pyhmmer")) or not os.path.isdir(libs):
        raise RuntimeError("no installed pyhmmer under %s" % site)
    out = os.path.join(HERE, "pyhmmer_cuda" + sysconfig.get_config_var("EXT_SUFFIX"))
    srcs = [os.path.join(HERE, f) for f in ("pyhmmer_cuda.pyx", "b2h_pyhmmer_glue.c", "b2h_pyhmmer_glue.h")] + [os.path.join(ROOT, "include", "b2h.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in srcs):
        return out
    cfile = os.path.join(HERE, "_pyhmmer_cuda.c")
    run = lambda cmd: subprocess.run(cmd, check=True)
    run([sys.executable, "-m", "cython", "-3", os.path.join(HERE, "pyhmmer_cuda.pyx"), "--output-file", cfile,
         "-I", os.path.join(libs, "cython", "include"), "-I", site, "-X", "cdivision=True", "-X", "nonecheck=False",
         "-E", "HMMER_IMPL=SSE", "-E", "LIMITED_API=True", "-E", "TARGET_SYSTEM=Linux", "-E", "TARGET_CPU=x86_64",
         "-E", "SSE2_BUILD_SUPPORT=True", "-E", "AVX2_BUILD_SUPPORT=False", "-E", "NEON_BUILD_SUPPORT=False", "-E", "MMX_BUILD_SUPPORT=False",
         "-E", "AVX512_BUILD_SUPPORT=False", "-E", "SYS_IMPLEMENTATION_NAME=cpython", "-E", "SYS_VERSION_INFO_MAJOR=%d" % sys.version_info[0],
         "-E", "SYS_VERSION_INFO_MINOR=%d" % sys.version_info[1], "-E", "SYS_BYTEORDER=little", "-E", "PYPY=False", "-E", "PROJECT_VERSION=0.12.3"])
    inc = [os.path.join(libs, "include"), os.path.join(libs, "include", "libeasel"), os.path.join(libs, "include", "libhmmer"),
           os.path.join(ROOT, "include"), HERE, sysconfig.get_paths()["include"]]
    cmd = ["gcc", "-shared", "-fPIC", "-O2", "-msse4.1", "-w", "-DPy_LIMITED_API=0x030C0000", "-DCYTHON_LIMITED_API=1", "-DCYTHON_USE_PYLONG_INTERNALS=0",
           cfile, os.path.join(HERE, "b2h_pyhmmer_glue.c"), "-o", out]
    for i in inc:
        cmd += ["-I", i]
    cmd += ["-L", libs, "-llibhmmer", "-llibeasel", "-L", PKG, "-lb2h", "-lm",
            "-Wl,-rpath," + libs, "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath,$ORIGIN/../../baseline/_ref/pyhmmer.libs"]
    run(cmd)
    return out


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else None, force="--force" in sys.argv))
