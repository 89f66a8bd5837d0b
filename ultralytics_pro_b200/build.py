"""In-tree build of libyolopost_b200.so (nvcc, sm_100a only).

The shared library lands next to the sources' package (``ultralytics_pro_b200/_lib/``) so it travels with the
repo snapshot; nothing is cached under ``~/.cache``.  ``python -m ultralytics_pro_b200.build`` rebuilds it.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libyolopost_b200.so")
STAMP = os.path.join(LIB_DIR, "libyolopost_b200.stamp")
SOURCES = ("ypb_decode.cu", "ypb_scan_tma.cu", "ypb_nms.cu", "ypb_post.cu", "ypb_abi.cu")
HEADERS = ("ypb_common.cuh", os.path.join("..", "..", "include", "yolopost_b200.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-shared", "-cudart", "shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libyolopost_b200 has no CPU fallback and cannot be built without the CUDA toolkit")


def _source_digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _source_digest()


def _compile_one(nvcc: str, src: str, obj: str, verbose: bool):
    cmd = [nvcc, *[f for f in NVCC_FLAGS if f not in ("-shared",)], *os.environ.get("YPB_EXTRA_NVCC", "").split()]
    # "-cudart shared" is a link-time option; harmless at compile time
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return src, cmd, proc


def _unit_digest(src: str) -> str:
    h = hashlib.sha256()
    for name in (src,) + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("YPB_EXTRA_NVCC", "").encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the translation units (in parallel, objects cached per unit under _lib/obj) and link one shared object.

    Safe under concurrent callers (torchrun ranks finding a stale library at the same time): the whole build runs under an
    exclusive file lock, and the library is linked to a temporary name and renamed into place, so no process can ever
    ``dlopen`` a partially written file."""
    if not force and is_current():
        return LIB_PATH
    import fcntl

    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():  # another process built it while we waited
                return LIB_PATH
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    from concurrent.futures import ThreadPoolExecutor

    nvcc = _nvcc()
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    jobs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        tag = obj + ".digest"
        objs.append(obj)
        digest = _unit_digest(src)
        fresh = os.path.exists(obj) and os.path.exists(tag) and open(tag).read().strip() == digest
        if force or verbose or not fresh:
            jobs.append((src, obj, tag, digest))
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as pool:
        results = list(pool.map(lambda j: _compile_one(nvcc, j[0], j[1] + ".tmp", verbose), jobs))
    for (src, obj, tag, digest), (_, cmd, proc) in zip(jobs, results):
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if verbose:
            sys.stderr.write(proc.stderr)
        os.replace(obj + ".tmp", obj)
        with open(tag, "w") as fh:
            fh.write(digest)
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", *objs, "-o", tmp]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    with open(STAMP + ".tmp", "w") as fh:
        fh.write(_source_digest())
    os.replace(STAMP + ".tmp", STAMP)
    return LIB_PATH


HARNESS_SRC = os.path.join(PKG_DIR, "..", "tests", "cabi_device_harness.cu")
HARNESS_PATH = os.path.join(LIB_DIR, "ypb_cabi_harness")


def build_harness(force: bool = False) -> str:
    """Compile tests/cabi_device_harness.cu - the torch-free C++ consumer of the C-ABI that the GPU tests and the
    compute-sanitizer runs execute - next to the library (rpath $ORIGIN).  Test infrastructure, not part of the product."""
    lib = build_library()
    src = os.path.normpath(HARNESS_SRC)
    h = hashlib.sha256()
    for name in (src, os.path.join(CSRC, HEADERS[1])):
        with open(name, "rb") as fh:
            h.update(fh.read())
    with open(STAMP) as fh:
        h.update(fh.read().encode())
    digest, tag = h.hexdigest(), HARNESS_PATH + ".digest"
    if not force and os.path.exists(HARNESS_PATH) and os.path.exists(tag) and open(tag).read().strip() == digest:
        return HARNESS_PATH
    import fcntl

    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            tmp = HARNESS_PATH + f".tmp{os.getpid()}"
            cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-lineinfo",
                   "-I", os.path.join(PKG_DIR, "..", "include"), src, "-o", tmp, "-L", os.path.dirname(lib), "-lyolopost_b200",
                   "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN", "-cudart", "shared"]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
            os.replace(tmp, HARNESS_PATH)
            with open(tag, "w") as fh:
                fh.write(digest)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return HARNESS_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_harness(force="--force" in sys.argv))
