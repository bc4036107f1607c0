#!/usr/bin/env python3
"""Build the parity checkers (TEST INFRASTRUCTURE ONLY - never on the product path).

Two shared libraries are produced:

* ``oracle/libfovoracle.so``   - the plain-C restatement ``oracle/fov_oracle.c``
  (``kind = "port"``).  Always buildable; travels to the GPU box.
* ``oracle/_ref/libfovref.so`` - the reference's *own* OpenCL kernel sources,
  compiled by g++ through ``oracle/ref_shim/clshim.h`` (``kind = "reference"``).
  Built only where ``/root/reference/src`` (or ``$FOV_REF_DIR``) exists, i.e. in
  the build container; the prebuilt .so travels to the GPU box (``oracle/_ref/``
  is git-ignored but not gpurun-ignored).  The reference sources are read where
  they lie; the only rewrite is the OpenCL vector-literal cast
  ``(int2)(a, b)`` -> ``int2(a, b)``, applied into a temp dir that is deleted
  after the compile.  No reference source is copied into this repository.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("FOV_REF_DIR", "/root/reference/src")

CL_FILES = [
    "sat_encoder_encode_kernels.cl",
    "sat_decoder_decode_kernel.cl",
    "sat_decoder_sample_rect_kernel.cl",
    "sat_decoder_interpolate_kernel.cl",
    "image_sampler_sample_rect_kernel.cl",
    "image_sampler_sample_logpolar_kernel.cl",
    "image_sampler_interpolate_kernel.cl",
    "projections_program.cl",
]

# -ffp-contract=off: mix() and the transform formulas must round every float
# operation separately (no FMA), exactly like scalar SSE2 code.
COMMON = ["-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math"]

VEC_LITERAL = re.compile(r"\((u?(?:char|short|int|float)[234])\)\(")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_port(force: bool = False) -> str:
    src = os.path.join(HERE, "fov_oracle.c")
    out = os.path.join(HERE, "libfovoracle.so")
    if not force and _newer(out, [src, __file__]):
        return out
    cmd = ["gcc", "-std=gnu11", *COMMON, "-o", out, src, "-lm"]
    subprocess.check_call(cmd)
    return out


def build_ref(force: bool = False) -> str | None:
    out_dir = os.path.join(HERE, "_ref")
    out = os.path.join(out_dir, "libfovref.so")
    shim = os.path.join(HERE, "ref_shim")
    entry = os.path.join(shim, "ref_entry.cc")
    if not os.path.isdir(REF_DIR):
        return out if os.path.exists(out) else None
    srcs = [os.path.join(REF_DIR, f) for f in CL_FILES] + [os.path.join(REF_DIR, "sat_encoder.cc")]
    if not force and _newer(out, srcs + [entry, os.path.join(shim, "clshim.h"), __file__]):
        return out
    os.makedirs(out_dir, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="fovref_")
    try:
        for f in CL_FILES:
            with open(os.path.join(REF_DIR, f), "r") as fh:
                text = fh.read()
            with open(os.path.join(tmp, f + ".inc"), "w") as fh:
                fh.write(VEC_LITERAL.sub(r"\1(", text))
        # SATEncoder::EncodeFrameCPU: the function's text, signature line to closing brace
        with open(os.path.join(REF_DIR, "sat_encoder.cc")) as fh:
            cc = fh.read().split("\n")
        a = next(i for i, ln in enumerate(cc) if ln.startswith("void SATEncoder::EncodeFrameCPU("))
        z = next(i for i in range(a + 1, len(cc)) if cc[i].startswith("}"))
        with open(os.path.join(tmp, "encode_frame_cpu.cc.inc"), "w") as fh:
            fh.write("\n".join(cc[a:z + 1]) + "\n")
        cmd = ["g++", "-std=c++17", *COMMON, "-w", "-I", shim, "-I", tmp, "-o", out, entry]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def main() -> None:
    force = "--force" in sys.argv
    print("port:", build_port(force))
    print("ref :", build_ref(force))


if __name__ == "__main__":
    main()
