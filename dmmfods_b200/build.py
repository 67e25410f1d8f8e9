"""In-tree build of libdmmfods_b200.so (hand-written sm_100a CUDA + the C-ABI).

`python -m dmmfods_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
Objects go to dmmfods_b200/_build/ (git-ignored); the .so sits next to the package so that it
travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.environ.get("DMM_BUILD_CSRC") or os.path.join(HERE, "csrc")      # experiment builds may compile another checkout's sources
BUILD = os.path.join(HERE, "_build")
# experiment builds (scripts/): DMM_BUILD_DEFINES="-DX -DY" + DMM_BUILD_TAG=name -> objects in _build_<name>/, libdmmfods_b200_<name>.so;
# the package loads such a library only when DMM_B200_LIB points at it
_TAG = os.environ.get("DMM_BUILD_TAG", "")
if _TAG:
    BUILD += "_" + _TAG
LIB = os.path.join(HERE, "libdmmfods_b200%s.so" % ("_" + _TAG if _TAG else ""))
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["common.cu", "igemm.cu", "igemm2.cu", "wgrad.cu", "elementwise.cu", "scatter.cu", "strict.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", INCLUDE] + os.environ.get("DMM_BUILD_DEFINES", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def _newer(target, deps):
    if not os.path.isfile(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_library(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(INCLUDE, "dmmfods_b200.h")]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if not force and _newer(o, [s] + headers):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out.decode("utf-8", "replace")))
    if force or procs or not _newer(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s" % r.stdout.decode("utf-8", "replace"))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
