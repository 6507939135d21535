"""The C-ABI library loads on a machine without a GPU, exports every symbol include/i3rc_b200.h declares, and
refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from i3rc_monte_carlo_model_b200 import _abi
from i3rc_monte_carlo_model_b200._lib import LIB_PATH, backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "i3rc_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(i3rc_[A-Za-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_reference_module_interface():
    names = declared_functions()
    # the public list of module monteCarloRadiativeTransfer (MCRT:154-156)
    for proc in ("new_Integrator", "copy_Integrator", "isReady_Integrator", "finalize_Integrator", "specifyParameters",
                 "computeRadiativeTransfer", "reportResults"):
        assert "i3rc_" + proc in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_struct_layouts_match_the_header():
    # sizes that the C compiler produces for the header's structs (LP64)
    assert C.sizeof(_abi.PhaseTable) == 48
    assert C.sizeof(_abi.Component) == 24 + 16 + 48
    assert C.sizeof(_abi.PhotonSource) == 64 + 40
    assert C.sizeof(_abi.Counters) == 14 * 8
    assert C.sizeof(_abi.Params) == 112


def test_no_cpu_fallback_without_a_device():
    be = backend()
    if be.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    import numpy as np
    e = np.array([0.0, 1.0], np.float32)
    one = np.ones(1, np.float32)
    pf = np.ones(1, np.int32)
    rc = be.new_Integrator(1, 1, 1, 1, _abi.fptr(e), _abi.fptr(e), _abi.fptr(e), _abi.fptr(one), _abi.fptr(one),
                           _abi.fptr(one), _abi.iptr(pf), C.byref(h))
    assert rc == _abi.FAILURE and not h.value
    assert b"no CUDA device" in be.last_message(None)
    assert be.version().startswith(b"i3rc_b200")


def test_product_package_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "i3rc_monte_carlo_model_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"i3rc_oracle", r"orc_[a-zA-Z]", r"oracle[/\\.]_?(build|binding)"):
                    assert not re.search(pat, text, flags=re.M), f"{f} uses the oracle ({pat})"
