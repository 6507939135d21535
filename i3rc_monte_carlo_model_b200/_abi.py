"""ctypes view of include/i3rc_b200.h.

The struct layouts here are the ones declared in ``include/i3rc_b200.h``.  ``Backend`` binds the entry
points of one shared library by prefix (``i3rc_`` for the product); the test suite reuses it for the CPU
checker, which deliberately exports the same layouts under another prefix.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

SUCCESS, WARNING, FAILURE = 0, 1, 2

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class PhaseTable(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_entries", C.c_int32),
        ("coef_offsets", c_int32_p),
        ("coefs", c_float_p),
        ("n_angles", C.c_int32),
        ("angles", c_float_p),
        ("values", c_float_p),
    ]


class Component(C.Structure):
    _fields_ = [
        ("extinction", c_float_p),
        ("ssa", c_float_p),
        ("phase_index", c_int32_p),
        ("horizontally_uniform", C.c_int32),
        ("z_level_base", C.c_int32),
        ("nz", C.c_int32),
        ("table", PhaseTable),
    ]


class Params(C.Structure):
    _fields_ = [
        ("present", C.c_uint32),
        ("surfaceAlbedo", C.c_float),
        ("minForwardTableSize", C.c_int32),
        ("minInverseTableSize", C.c_int32),
        ("numIntensityDirections", C.c_int32),
        ("intensityMus", c_float_p),
        ("intensityPhis", c_float_p),
        ("computeIntensity", C.c_int32),
        ("useRayTracing", C.c_int32),
        ("useRussianRoulette", C.c_int32),
        ("useRussianRouletteForIntensity", C.c_int32),
        ("zetaMin", C.c_float),
        ("useHybridPhaseFunsForIntenCalcs", C.c_int32),
        ("hybridPhaseFunWidth", C.c_float),
        ("numOrdersOrigPhaseFunIntenCalcs", C.c_int32),
        ("limitIntensityContributions", C.c_int32),
        ("maxIntensityContribution", C.c_float),
        ("surf_nx", C.c_int32),
        ("surf_ny", C.c_int32),
        ("surf_x", c_float_p),
        ("surf_y", c_float_p),
        ("surf_params", c_float_p),
    ]


P_BITS = {
    "surfaceAlbedo": 1 << 0,
    "surfaceBDRF": 1 << 1,
    "minForwardTableSize": 1 << 2,
    "minInverseTableSize": 1 << 3,
    "intensityMus": 1 << 4,
    "intensityPhis": 1 << 5,
    "computeIntensity": 1 << 6,
    "useRayTracing": 1 << 7,
    "useRussianRoulette": 1 << 8,
    "useRussianRouletteForIntensity": 1 << 9,
    "zetaMin": 1 << 10,
    "useHybridPhaseFunsForIntenCalcs": 1 << 11,
    "hybridPhaseFunWidth": 1 << 12,
    "numOrdersOrigPhaseFunIntenCalcs": 1 << 13,
    "limitIntensityContributions": 1 << 14,
    "maxIntensityContribution": 1 << 15,
}

SRC_DIRECTIONAL, SRC_RANDOM_AZIMUTH, SRC_FLUX, SRC_SPOTLIGHT = 1, 2, 3, 4
SRC_INTERNAL_FLUX, SRC_INTERNAL_INTENSITY, SRC_ARRAYS = 5, 6, 7


class PhotonSource(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("reserved", C.c_int32),
        ("numberOfPhotons", C.c_int64),
        ("solarMu", C.c_float),
        ("solarAzimuth", C.c_float),
        ("x", C.c_float),
        ("y", C.c_float),
        ("z", C.c_float),
        ("detectorMu", C.c_float),
        ("detectorPhi", C.c_float),
        ("detectorPointsUp", C.c_int32),
        ("has_deltaX", C.c_int32),
        ("has_deltaY", C.c_int32),
        ("deltaX", C.c_float),
        ("deltaY", C.c_float),
        ("xPosition", c_float_p),
        ("yPosition", c_float_p),
        ("zPosition", c_float_p),
        ("initialMu", c_float_p),
        ("initialPhi", c_float_p),
    ]


COUNTER_FIELDS = [
    "photons",
    "bad",
    "crossings_photon",
    "crossings_intensity",
    "collisions",
    "absorptions",
    "contributions",
    "exits_top",
    "surface_hits",
    "rng_draws",
    "roulette_kills",
    "null_collisions",
    "cells_skipped",
    "cells_skipped_intensity",
]


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in COUNTER_FIELDS]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n in COUNTER_FIELDS}


class StatsOut(C.Structure):
    _fields_ = [
        (n, c_double_p)
        for n in (
            "meanFluxUp",
            "meanFluxDown",
            "meanFluxAbsorbed",
            "fluxUp",
            "fluxDown",
            "fluxAbsorbed",
            "absorbedProfile",
            "absorbedVolume",
            "radiance",
            "meanRadiance",
        )
    ]


def f32(a):
    """Contiguous float32 copy/view in C order (callers pass arrays already in (z,y,x) memory order)."""
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fptr(a):
    return a.ctypes.data_as(c_float_p) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(c_int32_p) if a is not None else None


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


class Backend:
    """Entry points of one shared library, bound by prefix, with argument types set."""

    def __init__(self, lib: C.CDLL, prefix: str, name: str):
        self.lib, self.prefix, self.name = lib, prefix, name
        vp, ci = C.c_void_p, C.c_int
        sig = {
            "new_Integrator": (ci, [ci, ci, ci, ci, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_int32_p, C.POINTER(vp)]),
            "new_Integrator_components": (ci, [ci, ci, ci, c_float_p, c_float_p, c_float_p, ci, C.POINTER(Component), C.POINTER(vp)]),
            "set_phase_table": (ci, [vp, ci, C.POINTER(PhaseTable)]),
            "copy_Integrator": (ci, [vp, C.POINTER(vp)]),
            "finalize_Integrator": (None, [vp]),
            "isReady_Integrator": (ci, [vp]),
            "last_message": (C.c_char_p, [vp]),
            "specifyParameters": (ci, [vp, C.POINTER(Params)]),
            "computeRadiativeTransfer": (ci, [vp, C.POINTER(PhotonSource), c_int32_p, ci]),
            "reportResults": (ci, [vp] + [c_float_p] * 10),
            "get_intensityByComponent": (ci, [vp, c_float_p]),
            "get_counters": (None, [vp, C.POINTER(Counters)]),
            "tabulate": (ci, [vp]),
            "get_table": (ci, [vp, ci, ci, c_float_p, C.POINTER(ci), C.POINTER(ci)]),
            "trace_rays": (ci, [vp, ci, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_int32_p]),
            "sample_scattering_angles": (ci, [vp, ci, ci, ci, c_float_p, c_float_p]),
            "lookup_phase_function": (ci, [vp, ci, ci, ci, ci, c_float_p, c_float_p]),
        }
        self._optional = {
            "set_component_profile": (ci, [vp, ci, c_float_p]),
            "set_inverse_table": (ci, [vp, ci, ci, ci, c_float_p]),
            "set_forward_table": (ci, [vp, ci, ci, ci, c_float_p, c_float_p]),
            "stats_reset": (ci, [vp, ci]),
            "stats_accumulate": (ci, [vp]),
            "stats_device_buffer": (ci, [vp, C.POINTER(vp), C.POINTER(C.c_int64)]),
            "stats_report": (ci, [vp, C.c_double, ci, C.POINTER(StatsOut)]),
            "run_batches": (ci, [vp, C.POINTER(PhotonSource), C.c_int32, ci, ci, ci]),
            "comm_unique_id": (ci, [C.c_char_p]),
            "comm_init": (ci, [vp, ci, ci, C.c_char_p]),
            "stats_allreduce": (ci, [vp]),
            "comm_finalize": (ci, [vp]),
            "synchronize": (ci, [vp]),
            "stream": (vp, [vp]),
            "get_timing": (ci, [vp, c_double_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
            "reset_timing": (ci, [vp]),
            "set_tuning": (ci, [vp, C.c_char_p, ci]),
            "get_layout": (ci, [vp, ci]),
            "version": (C.c_char_p, []),
            "device_count": (ci, []),
            "set_device": (ci, [ci]),
            "probe_philox": (ci, [C.c_uint32, C.c_uint32, ci, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint32), c_float_p]),
            "probe_next_direct": (ci, [C.c_uint32, C.c_uint32, ci, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), c_float_p,
                                       c_float_p, c_float_p, C.POINTER(C.c_uint32)]),
            "measure_gather_rate": (ci, [C.c_size_t, ci, c_double_p]),
            "measure_issue_rate": (ci, [ci, c_double_p]),
        }
        for fn, (res, args) in sig.items():
            f = getattr(lib, prefix + fn)
            f.restype, f.argtypes = res, args
            setattr(self, fn, f)
        for fn, (res, args) in self._optional.items():
            f = getattr(lib, prefix + fn, None)
            if f is not None:
                f.restype, f.argtypes = res, args
            setattr(self, fn, f)
