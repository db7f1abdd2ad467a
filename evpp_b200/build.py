"""Build the CUDA shared library in-tree (evpp_b200/lib/libevpp_b200.so) with nvcc for sm_100a.

Every source is compiled to its own object (in parallel, re-compiled only when it or a header changed) and the
objects are linked into one shared library; the objects live next to the library and are git-ignored (and not
shipped to the GPU box: freshness is decided by a content hash stored next to the library, so a shipped library that
matches the shipped sources is never rebuilt there)."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libevpp_b200.so")
SOURCES = ["engine.cu", "sampling.cu", "mlp_simt.cu", "mlp_tc.cu", "tsdf.cu", "ngp.cu", "surrogate.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--split-compile", "0"] + os.environ.get("EVPP_NVCC_EXTRA", "").split()   # experiments: -D...


def _headers():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
        [os.path.join(ROOT, "..", "include", "evpp_b200.h")]


def _obj(src):
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _digest(paths):
    """Content hash (not mtimes: a snapshot copy of the tree need not preserve them) of files + compiler flags."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def _read(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


def build(force=False, verbose=False):
    hdrs = _headers()
    want = _digest(hdrs + [os.path.join(CSRC, s) for s in SOURCES])
    if not force and os.path.exists(LIB) and _read(LIB + ".hash") == want:
        return LIB                      # the library matches the sources: nothing to do (objects not needed)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    want_obj = {s: _digest(hdrs + [os.path.join(CSRC, s)]) for s in SOURCES}
    stale = [s for s in SOURCES if force or not os.path.exists(_obj(s)) or _read(_obj(s) + ".hash") != want_obj[s]]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", _obj(src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    if stale:
        with ThreadPoolExecutor(max_workers=max(1, min(len(stale), os.cpu_count() or 1))) as ex:
            for src, res in ex.map(compile_one, stale):
                if res.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
                open(_obj(src) + ".hash", "w").write(want_obj[src])
                if verbose:
                    print(res.stderr)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] +
                         [_obj(s) for s in SOURCES] + ["-ldl"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    open(LIB + ".hash", "w").write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
