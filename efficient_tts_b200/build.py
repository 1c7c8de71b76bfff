"""Build the sm_100a shared library in-tree with nvcc (no torch dependency in the binary).

``python -m efficient_tts_b200.build`` or ``build_library()``; the result,
``efficient_tts_b200/libefts_b200.so``, is git-ignored but travels with the source tree.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libefts_b200.so")
SOURCES = [os.path.join(CSRC, "efts_api.cu")]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
    os.path.join(os.path.dirname(HERE), "include", "efts_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def source_hash():
    """sha256 (first 16 hex digits) over the sources the library is built from; compiled into the library
    (``efts_version()``) and stamped into profiles/ so that a profile can be matched to the binary it describes."""
    import hashlib
    h = hashlib.sha256()
    for p in sorted(SOURCES + HEADERS):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the efts_b200 library cannot be built")
    return nvcc


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    """Compile ``csrc/*.cu`` for sm_100a into ``libefts_b200.so``; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + ['-DEFTS_SOURCE_SHA="%s"' % source_hash()] + \
        (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout)
    if verbose:
        print(r.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
