"""Builds libbq_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbq_b200.so")
SOURCES = ["bq_setup.cu", "bq_score.cu", "bq_reduce.cu", "bq_round.cu", "bq_capi.cu"]
HEADERS = ["bq_common.cuh", os.path.join("..", "..", "include", "bq_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        r = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(r.stdout)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stdout))
        objs.append(obj)
    r = subprocess.run([_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(os.path.join(CSRC, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
