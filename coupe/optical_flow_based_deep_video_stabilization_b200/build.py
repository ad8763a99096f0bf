"""In-tree build of libofstab.so (nvcc, sm_100a only).

    python -m coupe.optical_flow_based_deep_video_stabilization_b200.build [--force]

The library is a plain C-ABI shared object (no torch, no pybind): nvcc cross-compiles it on a
box without a GPU, and the built .so travels with the source tree to the GPU box.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
LIB_PATH = os.path.join(HERE, "libofstab.so")
STAMP = os.path.join(HERE, ".libofstab.stamp")

SOURCES = ["ofs_common.cu", "samplers.cu", "conv_gemm.cu", "flownet.cu", "clip.cu", "modes.cu", "host_pack.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libofstab.so cannot be built (there is no CPU fallback)")


def source_digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../../include/ofstab.h"]
    for name in files:
        path = os.path.normpath(os.path.join(CSRC, name))
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == source_digest()


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu (+ host_pack.cpp) into libofstab.so when sources changed.  Returns the library path.

    Safe under concurrent callers (torchrun ranks importing a stale tree at the same moment): one process builds
    under an exclusive file lock into a temporary file and renames it into place; the others wait for the lock,
    find the library fresh and return.  A reader never sees a half-written .so."""
    if not force and is_fresh():
        return LIB_PATH
    with open(os.path.join(HERE, ".libofstab.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_fresh():          # built by another process while this one waited
                return LIB_PATH
            digest = source_digest()
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = [_nvcc()] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                print(" ".join(cmd), flush=True)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            if os.path.exists(STAMP):
                os.remove(STAMP)                  # never a new library under an old stamp or vice versa
            os.replace(tmp, LIB_PATH)
            with open(STAMP + ".tmp", "w") as f:
                f.write(digest)
            os.replace(STAMP + ".tmp", STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
