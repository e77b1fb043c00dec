"""Build recipe for libw2e.so (explicit nvcc, sm_100a only, in-tree output).

    python -m where2edit_b200.build [--force]

The shared library has a plain C ABI (include/w2e.h) and links the CUDA runtime statically, so it
loads on a machine without a GPU driver (symbol checks on the CPU box) and travels to the GPU box
with the repo snapshot.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libw2e.so")
STAMP = os.path.join(PKG, "csrc", ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libw2e.so cannot be built")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _fingerprint():
    h = hashlib.sha256()
    files = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "w2e.h")]
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())   # relative: the snapshot on a GPU box lives under another root
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _fingerprint()


def build(force=False, verbose=False, jobs=None):
    """Compile every .cu under csrc/ to an object and link libw2e.so.  Returns the library path."""
    if not force and is_current():
        return LIB
    # one builder at a time: with one process per GPU (torchrun) every rank may find the library stale at once
    import fcntl
    lock = open(os.path.join(PKG, ".build_lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if is_current():            # another process built it while this one waited
            return LIB
        return _build_locked(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose=False):
    nvcc = _nvcc()
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            failed.append((src, out))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(f"--- {s}\n{o}" for s, o in failed))
    tmp = LIB + f".tmp{os.getpid()}"
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-o", tmp, *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout)
    os.replace(tmp, LIB)            # atomic: a concurrent loader sees the old or the new library, never a partial one
    with open(STAMP, "w") as fh:
        fh.write(_fingerprint())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
