"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports every symbol
that include/mogp_b200.h declares; status codes / enums agree between the header and the Python shim; calls
that need a device fail loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mogp_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mogp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = _declared_functions()
    for must in ["mogp_create", "mogp_destroy", "mogp_fit", "mogp_predict", "mogp_predict_allgather", "mogp_get",
                 "mogp_logpost_grad", "mogp_device_count", "mogp_last_error", "mogp_comm_create"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from mogp_emulator_b200 import libmogp
    assert libmogp.HAVE_LIBMOGP, "libmogp_b200.so must be built (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(libmogp.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), "missing export: " + name
    # and the shim binds exactly the declared set
    assert sorted(libmogp.SIGNATURES) == _declared_functions()


def test_enums_match_header():
    from mogp_emulator_b200 import libmogp
    text = open(HEADER).read()

    def enum_value(name):
        m = re.search(r"\b%s\s*=\s*(\d+)" % name, text)
        assert m, name
        return int(m.group(1))

    assert enum_value("MOGP_NUG_ADAPTIVE") == libmogp.nugget_type.adaptive == 0
    assert enum_value("MOGP_NUG_FIT") == libmogp.nugget_type.fit == 1
    assert enum_value("MOGP_NUG_FIXED") == libmogp.nugget_type.fixed == 2
    assert enum_value("MOGP_KERNEL_SQEXP") == libmogp.kernel_type.SquaredExponential
    assert enum_value("MOGP_KERNEL_MATERN52") == libmogp.kernel_type.Matern52
    for nm, val in [("MOGP_OK", libmogp.OK), ("MOGP_ERR_CUDA", libmogp.ERR_CUDA), ("MOGP_ERR_ARG", libmogp.ERR_ARG),
                    ("MOGP_ERR_NOT_PD", libmogp.ERR_NOT_PD), ("MOGP_ERR_NOT_FIT", libmogp.ERR_NOT_FIT),
                    ("MOGP_ERR_NCCL", libmogp.ERR_NCCL), ("MOGP_ERR_NOMEM", libmogp.ERR_NOMEM)]:
        assert enum_value(nm) == val


def test_no_cpu_fallback_without_device():
    """Without a B200 the product must refuse to run (RuntimeError like the reference,
    GaussianProcessGPU.py:223-229) -- never compute on the CPU."""
    import mogp_emulator_b200 as m
    if m.gpu_usable():
        pytest.skip("a GPU is visible here")
    X = np.random.default_rng(0).random((10, 2))
    y = np.zeros(10)
    with pytest.raises(RuntimeError):
        m.GaussianProcessGPU(X, y)
    with pytest.raises(RuntimeError):
        m.MultiOutputGP_GPU(X, np.zeros((2, 10)))
    from mogp_emulator_b200 import libmogp
    assert libmogp.device_count() == 0
    h = ctypes.c_void_p()
    st = libmogp.load().mogp_create(libmogp.dptr(X), 10, 2, libmogp.dptr(y), 1, 0, 2, 1e-6, 0, 0, ctypes.byref(h))
    assert st == libmogp.ERR_CUDA and "device" in libmogp.last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mogp_emulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "gp_oracle" not in src and "refstub" not in src, f
