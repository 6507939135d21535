"""ctypes binding of the CPU oracle (oracle/_build/libi3rc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by the product package.  The oracle exports the product's entry points
under the ``orc_`` prefix with identical struct layouts, so the host-side mirror classes of the
package can be pointed at it (``backend=oracle_backend()``) and the parity tests read the same on
both sides.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from i3rc_monte_carlo_model_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libi3rc_oracle.so")
_backend = None


def build(force=False):
    src = os.path.join(_HERE, "i3rc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class BatchStats(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("nd", C.c_int32), ("with_volume", C.c_int32)] + [
        (n, _abi.c_double_p) for n in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxUp", "fluxDown",
                                       "fluxAbsorbed", "absorbedProfile", "absorbedVolume", "radiance", "meanRadiance")]


_fast_backend = None


def _cpu_tag():
    """A short name for this machine's CPU (model + instruction-set flags): -march=native builds are per machine."""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        keep = [ln for ln in txt.splitlines() if ln.startswith(("model name", "flags"))][:2]
    except OSError:
        keep = []
    return hashlib.sha1("\n".join(keep).encode()).hexdigest()[:12]


def fast_oracle_backend() -> _abi.Backend:
    """The same C source built -O3 -march=native ON THIS MACHINE, for bench.py's CPU timing arm only (the parity oracle is
    the -O2 -ffp-contract=off build).  Returns a backend with the batch driver bound."""
    global _fast_backend
    if _fast_backend is None:
        d = os.path.join(_HERE, "_build", "fast_" + _cpu_tag())
        subprocess.check_call(["make", "-C", _HERE, "-s", "fast", "FASTDIR=" + d], stdout=subprocess.DEVNULL)
        _fast_backend = _bind(C.CDLL(os.path.join(d, "libi3rc_oracle_fast.so")), "oracle-O3-native")
    return _fast_backend


def oracle_backend() -> _abi.Backend:
    global _backend
    if _backend is None:
        build()
        _backend = _bind(C.CDLL(LIB_PATH), "oracle")
    return _backend


def _bind(lib, name) -> _abi.Backend:
    be = _abi.Backend(lib, "orc_", name)
    vp, ci = C.c_void_p, C.c_int
    u32p = C.POINTER(C.c_uint32)
    lib.orc_mt_seed_vector.argtypes = [u32p, _abi.c_int32_p, ci]
    lib.orc_mt_seed_scalar.argtypes = [u32p, C.c_int32]
    lib.orc_mt_int32.argtypes, lib.orc_mt_int32.restype = [u32p], C.c_uint32
    lib.orc_mt_real.argtypes, lib.orc_mt_real.restype = [u32p], C.c_float
    lib.orc_findIndex.argtypes, lib.orc_findIndex.restype = [C.c_float, _abi.c_float_p, ci, ci], ci
    lib.orc_computeLobattoMus.argtypes = [_abi.c_float_p, ci]
    lib.orc_next_direct.argtypes = [_abi.c_float_p, ci, C.c_float, _abi.c_float_p]
    lib.orc_hybrid_phase_functions.argtypes = [_abi.c_float_p, ci, ci, _abi.c_float_p, C.c_float, _abi.c_float_p]
    lib.orc_run_batches.argtypes = [vp, C.POINTER(_abi.PhotonSource), C.c_int32, ci, ci, ci, ci,
                                    C.POINTER(BatchStats), C.POINTER(_abi.Counters)]
    lib.orc_run_batches.restype = ci
    return be


class MT19937:
    """RandomNumbersForMC.f95 restated (oracle) -- for the RNG pin tests."""

    def __init__(self, seed):
        self.lib = oracle_backend().lib
        self.state = (C.c_uint32 * 625)()
        seed = np.atleast_1d(np.asarray(seed, np.int32))
        if seed.size == 1:
            self.lib.orc_mt_seed_scalar(self.state, int(seed[0]))
        else:
            self.lib.orc_mt_seed_vector(self.state, _abi.iptr(seed), seed.size)

    def int32(self):
        return int(self.lib.orc_mt_int32(self.state))

    def real(self):
        return float(self.lib.orc_mt_real(self.state))


def run_batches(integ, photons, iseed, numBatches, seedOrder=0, batchBegin=1, nThreads=0, with_volume=False):
    """monteCarloDriver.f95:264-378 on the oracle with OpenMP threads standing in for MPI ranks.
    Returns (sums dict of float64 arrays shaped [2, ...], counters dict)."""
    be = integ.backend  # (the parity oracle or its timing build: whichever made the integrator)
    nx, ny, nz, nd = integ.nx, integ.ny, integ.nz, integ.nDir
    st = BatchStats(nx=nx, ny=ny, nz=nz, nd=nd, with_volume=int(with_volume))
    arrs = {
        "meanFluxUp": np.zeros(2), "meanFluxDown": np.zeros(2), "meanFluxAbsorbed": np.zeros(2),
        "fluxUp": np.zeros((2, ny, nx)), "fluxDown": np.zeros((2, ny, nx)), "fluxAbsorbed": np.zeros((2, ny, nx)),
        "absorbedProfile": np.zeros((2, nz)),
    }
    if with_volume:
        arrs["absorbedVolume"] = np.zeros((2, nz, ny, nx))
    if nd:
        arrs["radiance"] = np.zeros((2, nd, ny, nx))
        arrs["meanRadiance"] = np.zeros((2, nd))
    for k, a in arrs.items():
        setattr(st, k, _abi.dptr(a))
    cnt = _abi.Counters()
    src = photons.as_c()
    rc = be.lib.orc_run_batches(integ.handle, C.byref(src), int(iseed), int(seedOrder), int(batchBegin),
                                int(numBatches), int(nThreads), C.byref(st), C.byref(cnt))
    if rc == _abi.FAILURE:
        raise RuntimeError("oracle run_batches failed: " + integ._msg())
    return arrs, cnt.as_dict()
