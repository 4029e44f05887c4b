"""TEST INFRASTRUCTURE ONLY -- builds the oracle's C restatement of the reference's weave kernels
(oracle/weave_kernels.c) into oracle/_build/libgpo_weave.so with gcc -O3 -fopenmp.

oracle/_ref/ (a build of the reference's own sources) does not exist for this reference: GPclust is pure
Python whose only native code is C++ text inside Python strings handed to scipy.weave (Python 2 only),
and every module imports GPy at top level (absent, no network) -- "unbuildable", see DESIGN.md."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "weave_kernels.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libgpo_weave.so")


def _candidates():
    ccs = []
    for cc in (os.environ.get("CC"), "/usr/bin/gcc", "gcc", "cc"):
        if cc and cc not in ccs:
            ccs.append(cc)
    for cc in ccs:
        yield [cc, "-O3", "-fopenmp"], True
    for cc in ccs:  # a gcc wrapper that cannot find libgomp.spec: point it at the distro's gcc directory
        yield [cc, "-O3", "-fopenmp", "-B/usr/lib/gcc/x86_64-linux-gnu/13"], True
    for cc in ccs:
        yield [cc, "-O3"], False  # last resort: serial (bench.py then reports cores = 1)


def build_oracle(force=False):
    """-> path of the shared library.  oracle/_build/openmp.txt records whether OpenMP was linked."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    last = None
    for head, omp in _candidates():
        res = subprocess.run(head + ["-shared", "-fPIC", "-o", LIB, SRC, "-lm"], capture_output=True, text=True)
        if res.returncode == 0:
            with open(os.path.join(OUT_DIR, "openmp.txt"), "w") as f:
                f.write("1" if omp else "0")
            return LIB
        last = res.stderr
    raise RuntimeError("could not build oracle/weave_kernels.c: %s" % last)


def has_openmp():
    try:
        return open(os.path.join(OUT_DIR, "openmp.txt")).read().strip() == "1"
    except OSError:
        return False


if __name__ == "__main__":
    print(build_oracle(force=True))
