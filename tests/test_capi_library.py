"""
The C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol include/gpso_b200.h declares.
No compute entry point is called here (this file runs on the CPU-only build box).
"""
import ctypes
import os
import re

import numpy as np
import pytest

from pygpso_b200 import _build, backend

HEADER = os.path.join(os.path.dirname(__file__), "..", "include", "gpso_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    if _build.needs_build():
        _build.build_library()
    return backend.library_path()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpso_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert declared_symbols() == sorted(backend.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in gpso_b200.h but not exported"
    lib.gpso_version.restype = ctypes.c_int
    assert lib.gpso_version() == 1
    lib.gpso_grow_count.restype = ctypes.c_int64
    assert lib.gpso_grow_count(5) == 121 and lib.gpso_grow_count(12) == 265720


def test_argument_errors_do_not_need_a_gpu(lib_path):
    lib = backend.load_library()
    assert lib.gpso_create(0, 99, 0, 1, ctypes.byref(ctypes.c_void_p())) == -1  # GPSO_E_BADARG: unknown kernel
    assert b"kernel" in lib.gpso_last_error()
    assert lib.gpso_destroy(None) == 0


def test_library_is_sm100a_with_dmma(lib_path):
    """The built image carries sm_100a SASS with FP64 tensor-core instructions (DMMA) and async copies (LDGSTS)."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert sass.count("DMMA.8x8x4") > 500
    assert "LDGSTS" in sass


@pytest.mark.skipif(backend.load_library().gpso_device_count() > 0 if os.path.exists(backend.library_path()) else False,
                    reason="a GPU is present; this test checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu(lib_path):
    """No CPU fallback: fitting through the public API without a device raises instead of computing on the host."""
    from pygpso_b200 import GPRSurrogate

    backend.reset_default_backend()
    surr = GPRSurrogate.default()
    x = np.random.default_rng(0).random((5, 2))
    with pytest.raises(backend.GpsoBackendError):
        surr._gp_train(x, x[:, :1])
    with pytest.raises(backend.GpsoBackendError):
        from pygpso_b200 import ParameterSpace

        ParameterSpace([[0, 1], [0, 1]], ["a", "b"]).grow(3)
