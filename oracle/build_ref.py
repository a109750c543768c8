#!/usr/bin/env python
"""TEST INFRASTRUCTURE — builds the UNMODIFIED reference into oracle/_ref/.

Recipe; outputs go only to oracle/_ref/ (git-ignored, shipped by gpurun):

* the four Cython modules are cythonized *from where they lie* under
  /root/reference/bayesian_quadrature/*.pyx (language level 2, as the reference is
  Python-2 code) and compiled with gcc against two shim headers (oracle/shim/) that map
  the ATLAS names of linalg_c.pyx:14-45 onto scipy's bundled OpenBLAS/LAPACK;
* the pure-Python half (bq.py, util.py, __init__.py) is emitted with four mechanical
  py2->py3 touches that change no arithmetic: relative import in __init__.py:8,
  dict.iteritems -> items (bq.py:935,961), xrange -> range (bq.py:400,421,443,640;
  util.py:154), logger.warn kept;
* the un-vendored ``gp`` dependency is provided by oracle/gp_standin.py and a stub
  ``matplotlib.pyplot`` (bq.py:3 / util.py:1 import it at module level).

Nothing here is imported by the product package.  On a box without /root/reference
(the GPU box) this script is a no-op and the prebuilt oracle/_ref/ is used.
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BQ_REFERENCE_DIR", "/root/reference")
DEST = os.path.join(HERE, "_ref")
MODS = ["linalg_c", "gauss_c", "bq_c", "util_c"]


def ref_available():
    return os.path.isdir(os.path.join(REF, "bayesian_quadrature"))


def built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(DEST, "bayesian_quadrature", m + suffix)) for m in MODS) \
        and os.path.exists(os.path.join(DEST, "bayesian_quadrature", "bq.py")) \
        and os.path.exists(os.path.join(DEST, "gp.py"))


def _openblas():
    import scipy
    libs = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    cands = sorted(glob.glob(os.path.join(libs, "libscipy_openblas-*.so")))
    if not cands:
        raise RuntimeError("scipy's bundled LP64 OpenBLAS not found under %s" % libs)
    return cands[0], libs


def _py3(src, name):
    if name == "__init__.py":
        src = src.replace("from bq import BQ", "from .bq import BQ")
    src = src.replace(".iteritems()", ".items()")
    src = re.sub(r"\bxrange\(", "range(", src)
    return src


def build(force=False, verbose=False):
    if not ref_available():
        return built()
    if built() and not force:
        return True
    import numpy as np
    pkg = os.path.join(DEST, "bayesian_quadrature")
    bld = os.path.join(DEST, "_build")
    for d in (pkg, bld, os.path.join(DEST, "matplotlib")):
        os.makedirs(d, exist_ok=True)
    blas, blasdir = _openblas()
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    pyinc = sysconfig.get_paths()["include"]
    srcdir = os.path.join(REF, "bayesian_quadrature")
    for m in MODS:
        c = os.path.join(bld, m + ".c")
        subprocess.check_call(
            [sys.executable, "-m", "cython", "-2", "-I", srcdir, os.path.join(srcdir, m + ".pyx"), "-o", c],
            stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.DEVNULL)
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", os.path.join(HERE, "shim"), "-I", pyinc,
             "-I", np.get_include(), c, "-o", os.path.join(pkg, m + suffix), blas,
             "-Wl,-rpath," + blasdir, "-lm"])
    for name in ("__init__.py", "bq.py", "util.py"):
        with open(os.path.join(srcdir, name)) as fh:
            src = fh.read()
        with open(os.path.join(pkg, name), "w") as fh:
            fh.write(_py3(src, name))
    shutil.copyfile(os.path.join(HERE, "gp_standin.py"), os.path.join(DEST, "gp.py"))
    with open(os.path.join(DEST, "matplotlib", "__init__.py"), "w") as fh:
        fh.write("# stub: the reference imports matplotlib.pyplot at module level (bq.py:3)\n")
    with open(os.path.join(DEST, "matplotlib", "pyplot.py"), "w") as fh:
        fh.write("# stub\n")
    shutil.rmtree(bld, ignore_errors=True)
    return built()


def import_reference():
    """Return the reference package (``bayesian_quadrature``) and the ``gp`` stand-in,
    imported from oracle/_ref/.  Raises ImportError if oracle/_ref/ was never built."""
    if not built():
        raise ImportError("oracle/_ref/ is not built (run oracle/build_ref.py where /root/reference exists)")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import bayesian_quadrature
    import gp
    return bayesian_quadrature, gp


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
