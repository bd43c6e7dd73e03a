"""Build libklnmf.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m multimodal_b200.build [--force]

The shared object lands in multimodal_b200/csrc/libklnmf.so (git-ignored, but it travels
to the GPU box with the repo snapshot).  No JIT cache, no torch cpp_extension: the
boundary is a plain C ABI (include/klnmf.h) loaded with ctypes.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libklnmf.so")
SOURCES = ["api.cu", "dense_generic.cu", "dense_tc.cu", "dense_fused.cu", "dense_fused256.cu", "elementwise.cu", "sparse.cu", "evaluation.cu", "nccl_dyn.cu", "stack.cu"]
HEADERS = ["common.cuh", "tc_ptx.cuh", os.path.join("..", "..", "include", "klnmf.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "128,177"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libklnmf cannot be built")


def have_nvcc():
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


STAMP = LIB + ".srchash"


def source_hash():
    """sha256 over the compiler flags and every source / header: what the library was built from.  (Modification times
    do not survive a copy of the tree to another machine; the contents do.)"""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for rel in SOURCES + HEADERS:
        with open(os.path.join(CSRC, rel), "rb") as fh:
            h.update(rel.encode() + b"\0" + fh.read())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    try:
        with open(STAMP) as fh:
            return fh.read().strip() != source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libklnmf.so. Returns its path.  Safe to call from several
    processes at once (the ranks of a torchrun job): one builds under a file lock, the others find the result."""
    if not force and not is_stale():
        return LIB
    import fcntl
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    with open(os.path.join(CSRC, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():          # somebody else built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    digest = source_hash()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
        if verbose and r.stderr.strip():
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = "%s.tmp.%d" % (LIB, os.getpid())
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-ldl", "-Wno-deprecated-gpu-targets"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stderr[-4000:])
    os.replace(tmp, LIB)
    with open(STAMP + ".tmp", "w") as fh:
        fh.write(digest + "\n")
    os.replace(STAMP + ".tmp", STAMP)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
