"""CPU-side checks of the drop-in boundary: libfov360.so builds/loads, exports every symbol that
include/fov360.h declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import pytest


def test_library_exports_every_declared_symbol(fov):
    lib = fov.load()
    declared = fov.header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    # the Python prototypes cover the header exactly (nothing undeclared is bound)
    assert sorted(fov.PROTOTYPES) == declared


def test_library_has_sm100a_code_only(fov):
    out = subprocess.run(["cuobjdump", "-lelf", fov.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = {ln.split(".")[-2] for ln in out.stdout.splitlines() if ln.strip().endswith(".cubin")}
    assert archs == {"sm_100a"}, archs


def test_no_oracle_linked_into_product(fov):
    out = subprocess.run(["nm", "-D", fov.library_path()], capture_output=True, text=True).stdout
    assert "orc_" not in out and "ref_sat" not in out
    src_dir = os.path.join(os.path.dirname(fov.library_path()))
    for root, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "fov_oracle" not in text and "libfovoracle" not in text, f


def test_null_context_is_rejected(fov):
    lib = fov.load()
    assert lib.fov_sync(None) == -1
    assert lib.fov_sat_encode(None, None, None, 8, 8, 32) == -1
    assert b"not initialized" in lib.fov_last_error_string(None).lower()
    assert lib.fov_ctx_launch_count(None) == 0
    assert lib.fov_ctx_stream(None) is None


def test_reduced_dim_rule(fov):
    # parameters.h:8-9 and run_satlogrectilinear.cc:113-114
    assert (fov.reduced_dim(1920), fov.reduced_dim(1080)) == (1072, 608)
    assert (fov.reduced_dim(3840), fov.reduced_dim(1920)) == (2144, 1072)
    assert (fov.reduced_dim(7680), fov.reduced_dim(3840)) == (4272, 2144)
    assert (fov.REDUCED_BUFFER_WIDTH, fov.REDUCED_BUFFER_HEIGHT) == (1072, 608)


def test_context_creation_fails_loudly_without_gpu(fov):
    lib = fov.load()
    if lib.fov_device_count() > 0:
        pytest.skip("a GPU is present")
    err = C.c_int(0)
    assert not lib.fov_ctx_create(0, C.byref(err))
    assert err.value == -3
    assert b"no CPU fallback" in lib.fov_last_error_string(None)
    with pytest.raises(fov.FovError):
        fov.OpenCLManager(0).InitializeContext()
    with pytest.raises(fov.FovError):
        fov.SATEncoder().EncodeFrameGPU(0, 0, 8, 8, 32)  # "Not initialized with OpenCL"
