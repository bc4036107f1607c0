#!/usr/bin/env python3
"""In-tree nvcc build of libfov360.so (sm_100a only).

The shared library lands next to this file so that it travels to the GPU box with the
repository snapshot.  No JIT, no torch extension machinery: plain ``nvcc`` per translation
unit (run in parallel), then one link step.  ``luts.cc`` holds the host-side table builders
and is compiled with ``-ffp-contract=off`` because its truncating float formulas must round
every operation separately (see the header of that file).
"""
from __future__ import annotations

import concurrent.futures as cf
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libfov360.so")

SOURCES = ["capi.cu", "sat_encode.cu", "sat_onepass.cu", "sat_decode.cu", "image_sampler.cu",
           "projections.cu", "color_convert.cu", "luts.cc"]
HEADERS = [os.path.join(CSRC, "fov360_internal.h"), os.path.join(CSRC, "sat_common.cuh"), os.path.join(CSRC, "bounds_check.cuh"),
           os.path.join(CSRC, "projection_common.cuh"), os.path.join(CSRC, "pixel_math.cuh"),
           os.path.join(INCLUDE, "fov360.h")]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo", *ARCH,
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
    "-I", INCLUDE, "-I", CSRC,
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfov360.so cannot be built (there is no CPU fallback)")


MANIFEST = os.path.join(HERE, "libfov360.manifest")


def _digest(paths: list[str]) -> str:
    """Content hash of the build inputs: mtimes do not survive the snapshot to the GPU box."""
    # flags without the checkout-dependent absolute include paths
    h = hashlib.sha256(" ".join(f for f in NVCC_FLAGS if not f.startswith(ROOT)).encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(p.rsplit(os.sep, 1)[-1].encode() + b"\0" + fh.read())
    return h.hexdigest()


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(nvcc: str, src: str, obj: str, verbose: bool) -> None:
    cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed for " + src)


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds (if stale) and returns the path of libfov360.so."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    digest = _digest(srcs + HEADERS)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(MANIFEST):
        with open(MANIFEST) as fh:
            if fh.read().strip() == digest:
                return LIB_PATH
    nvcc = nvcc_path()
    os.makedirs(OBJ_DIR, exist_ok=True)
    # one builder at a time (torchrun starts one process per GPU against the same tree)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and os.path.exists(LIB_PATH) and os.path.exists(MANIFEST):
            with open(MANIFEST) as fh:
                if fh.read().strip() == digest:
                    return LIB_PATH  # another process finished the build while we waited
        return _build_locked(nvcc, srcs, digest, force, verbose)


def _build_locked(nvcc: str, srcs: list[str], digest: str, force: bool, verbose: bool) -> str:
    objs = [os.path.join(OBJ_DIR, os.path.splitext(s)[0] + ".o") for s in SOURCES]
    todo = [(s, o) for s, o in zip(srcs, objs)
            if force or _stale(o, [s] + HEADERS + [os.path.abspath(__file__)])]
    todo = todo if os.path.exists(MANIFEST) else list(zip(srcs, objs))
    with cf.ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for fut in [ex.submit(_compile, nvcc, s, o, verbose) for s, o in todo]:
            fut.result()
    tmp = LIB_PATH + ".tmp%d" % os.getpid()
    subprocess.check_call([nvcc, *ARCH, "-shared", "-cudart", "static", "-o", tmp, *objs])
    os.replace(tmp, LIB_PATH)
    with open(MANIFEST + ".tmp", "w") as fh:
        fh.write(digest + "\n")
    os.replace(MANIFEST + ".tmp", MANIFEST)
    return LIB_PATH


CHECK_LIB_PATH = os.path.join(HERE, "libfov360_check.so")
CHECK_MANIFEST = os.path.join(HERE, "libfov360_check.manifest")


def build_check(force: bool = False) -> str:
    """The same library with -DFOV360_BOUNDS_CHECK (csrc/bounds_check.cuh): test infrastructure for
    tests/test_gpu_bounds.py, never loaded unless FOV360_LIB points at it."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    digest = _digest(srcs + HEADERS) + "+check"
    if not force and os.path.exists(CHECK_LIB_PATH) and os.path.exists(CHECK_MANIFEST):
        with open(CHECK_MANIFEST) as fh:
            if fh.read().strip() == digest:
                return CHECK_LIB_PATH
    nvcc = nvcc_path()
    obj_dir = os.path.join(OBJ_DIR, "check")
    os.makedirs(obj_dir, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        objs = [os.path.join(obj_dir, os.path.splitext(s)[0] + ".o") for s in SOURCES]

        def one(src, obj):
            r = subprocess.run([nvcc, *NVCC_FLAGS, "-DFOV360_BOUNDS_CHECK", "-c", src, "-o", obj],
                               capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed for " + src)

        with cf.ThreadPoolExecutor(max_workers=max(1, min(len(srcs), os.cpu_count() or 1))) as ex:
            for fut in [ex.submit(one, s, o) for s, o in zip(srcs, objs)]:
                fut.result()
        tmp = CHECK_LIB_PATH + ".tmp%d" % os.getpid()
        subprocess.check_call([nvcc, *ARCH, "-shared", "-cudart", "static", "-o", tmp, *objs])
        os.replace(tmp, CHECK_LIB_PATH)
        with open(CHECK_MANIFEST, "w") as fh:
            fh.write(digest + "\n")
    return CHECK_LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
