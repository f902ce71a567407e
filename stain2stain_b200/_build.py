"""In-tree build of the sm_100a C-ABI library (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libs2s_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["common.cuh", "attention.cuh", "conv_igemm.cuh", "elementwise.cuh", "head_conv.cuh", "linear.cuh", "optim.cuh", "multitask.cuh", "tiles.cuh",
           os.path.join("..", "..", "include", "s2s_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library in-tree.  Safe when several ranks of one node call it at once (torchrun): one process builds
    under a file lock, into a temporary file that is renamed into place; the others wait and find it fresh."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another rank built it while we waited
                return LIB_PATH
            tmp = LIB_PATH + f".tmp.{os.getpid()}"
            cmd = [_nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                   "-shared", "-Xcompiler", "-fPIC", "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd), file=sys.stderr)
            try:
                subprocess.run(cmd, check=True, cwd=CSRC)
                os.replace(tmp, LIB_PATH)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
