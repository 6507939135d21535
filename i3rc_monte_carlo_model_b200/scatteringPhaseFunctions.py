"""Phase-function objects at the boundary (Code/scatteringPhaseFunctions.f95).

Only what the integrator path touches is mirrored: the two constructors of ``phaseFunction``
(:102, :164), the two of ``phaseFunctionTable`` (:227, :339), ``getInfo_*`` and the conversion to the
C-ABI ``i3rc_phase_table``.  Evaluation (``getPhaseFunctionValues``, :446-648) and normalisation
(:1329) happen on the device when the integrator tabulates its tables.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .ErrorMessages import setStateToFailure, setStateToSuccess, stateIsFailure

Pi = np.float32(3.141592654)
maxScatteringAngle = Pi


class phaseFunction:
    def __init__(self):
        self.scatteringAngle = None
        self.value = None
        self.legendreCoefficients = None
        self.extinction = 0.0
        self.singleScatteringAlbedo = 0.0
        self.description = ""


def _spacing(x):
    return np.spacing(np.float32(abs(x))) if x != 0 else np.finfo(np.float32).tiny


def new_PhaseFunction(*args, extinction=None, singleScatteringAlbedo=None, description="", status=None,
                      legendreCoefficients=None, scatteringAngle=None, value=None):
    """new_PhaseFunction(legendreCoefficients) or new_PhaseFunction(scatteringAngle, value)."""
    if len(args) == 1:
        legendreCoefficients = args[0]
    elif len(args) == 2:
        scatteringAngle, value = args
    p = phaseFunction()
    if extinction is not None and extinction < 0:
        setStateToFailure(status, "newPhaseFunction: negative extinction supplied.")
    if singleScatteringAlbedo is not None and not (0.0 <= singleScatteringAlbedo <= 1.0):
        setStateToFailure(status, "newPhaseFunction: singleScatteringAlbedo out of bounds.")
    if legendreCoefficients is not None:
        c = np.asarray(legendreCoefficients, dtype=np.float32).ravel()
        if c.size > 1 and (c[0] > 1.0 or c[0] < -1.0):
            setStateToFailure(status, "newPhaseFunction: Asymmetery parameter out of bounds.")
        if status is not None and stateIsFailure(status):
            return p
        p.legendreCoefficients = c.copy()
    else:
        a = np.asarray(scatteringAngle, dtype=np.float32).ravel()
        v = np.asarray(value, dtype=np.float32).ravel()
        _check_angles(a, v[:, None] if v.size == a.size else v, "newPhaseFunction", status)
        if status is not None and stateIsFailure(status):
            return p
        p.scatteringAngle, p.value = a.copy(), v.copy()
    if extinction is not None:
        p.extinction = float(extinction)
    if singleScatteringAlbedo is not None:
        p.singleScatteringAlbedo = float(singleScatteringAlbedo)
    p.description = description[:64]
    setStateToSuccess(status)
    return p


def _check_angles(a, v, who, status):
    if np.any(a < 0) or np.any(a > maxScatteringAngle):
        setStateToFailure(status, f"{who}: ScatteringAngle out of bounds.")
    if abs(a[0]) > _spacing(0.0):
        setStateToFailure(status, f"{who}: First scattering angle must be min value")
    if abs(a[-1] - maxScatteringAngle) > _spacing(maxScatteringAngle):
        setStateToFailure(status, f"{who}: Last scattering angle must be max value")
    if np.any(np.diff(a) <= 0):
        setStateToFailure(status, f"{who}: Scattering angle must be increasing, unique.")
    if np.any(v < 0):
        setStateToFailure(status, f"{who}: Negative phase function values supplied.")
    if v.shape[0] != a.size:
        setStateToFailure(status, f"{who}: Number of scattering angles and phase function values must match.")


class phaseFunctionTable:
    def __init__(self):
        self.phaseFunctions: list[phaseFunction] = []
        self.key = None
        self.description = ""
        self.oneAngleSet = False
        self._keep = None  # arrays referenced by the last as_c()

    def as_c(self) -> _abi.PhaseTable:
        """The i3rc_phase_table view (include/i3rc_b200.h)."""
        t = _abi.PhaseTable()
        n = len(self.phaseFunctions)
        t.n_entries = n
        if all(p.legendreCoefficients is not None for p in self.phaseFunctions):
            offs = np.zeros(n + 1, np.int32)
            offs[1:] = np.cumsum([p.legendreCoefficients.size for p in self.phaseFunctions])
            coefs = (np.concatenate([p.legendreCoefficients for p in self.phaseFunctions])
                     if offs[-1] > 0 else np.zeros(1, np.float32)).astype(np.float32)
            t.kind, t.coef_offsets, t.coefs = 1, _abi.iptr(offs), _abi.fptr(coefs)
            self._keep = (offs, coefs)
        elif self.oneAngleSet:
            ang = _abi.f32(self.phaseFunctions[0].scatteringAngle)
            vals = _abi.f32(np.stack([p.value for p in self.phaseFunctions]))
            t.kind, t.n_angles, t.angles, t.values = 2, ang.size, _abi.fptr(ang), _abi.fptr(vals)
            self._keep = (ang, vals)
        else:
            raise NotImplementedError(
                "phase function tables mixing angle sets are outside the scope of the C ABI "
                "(the reference cannot persist them either, scatteringPhaseFunctions.f95:905-907)")
        return t


def new_PhaseFunctionTable(*args, key=None, phaseFunctionDescriptions=None, tableDescription="", status=None,
                           extinction=None, singleScatteringAlbedo=None):
    """new_PhaseFunctionTable(phaseFunctions, key) or new_PhaseFunctionTable(scatteringAngle, values, key)."""
    table = phaseFunctionTable()
    if len(args) >= 1 and isinstance(args[0], (list, tuple)) and args[0] and isinstance(args[0][0], phaseFunction):
        pfs = list(args[0])
        if len(args) > 1:
            key = args[1]
        key = np.asarray(key, dtype=np.float32).ravel()
        if key.size != len(pfs):
            setStateToFailure(status, "newPhaseFunctionTable: Number of phase functions and key values must match.")
        if np.any(np.diff(key) <= 0):
            setStateToFailure(status, "newPhaseFunctionTable: Key values must be unique, increasing.")
        if status is not None and stateIsFailure(status):
            return table
        table.phaseFunctions, table.key, table.oneAngleSet = pfs, key, False
    else:
        a = np.asarray(args[0], dtype=np.float32).ravel()
        v = np.asarray(args[1], dtype=np.float32)  # (nAngles, nEntries) like the reference
        if len(args) > 2:
            key = args[2]
        key = np.asarray(key, dtype=np.float32).ravel()
        _check_angles(a, v, "newPhaseFunctionTable", status)
        if key.size != v.shape[1]:
            setStateToFailure(status, "newPhaseFunctionTable: Number of phase functions and key values must match.")
        if np.any(np.diff(key) <= 0):
            setStateToFailure(status, "newPhaseFunctionTable: Key values must be unique, increasing.")
        if status is not None and stateIsFailure(status):
            return table
        for i in range(v.shape[1]):
            p = phaseFunction()
            p.scatteringAngle, p.value = a, np.ascontiguousarray(v[:, i])
            if extinction is not None:
                p.extinction = float(extinction[i])
            if singleScatteringAlbedo is not None:
                p.singleScatteringAlbedo = float(singleScatteringAlbedo[i])
            table.phaseFunctions.append(p)
        table.key, table.oneAngleSet = key, True
    table.description = tableDescription[:1024]
    setStateToSuccess(status)
    return table


def getInfo_PhaseFunctionTable(table, status=None):
    setStateToSuccess(status)
    return {"nEntries": len(table.phaseFunctions), "key": table.key, "tableDescription": table.description}


def isReady_PhaseFunctionTable(table):
    return bool(table.phaseFunctions)


def finalize_PhaseFunctionTable(table):
    table.phaseFunctions, table.key, table.oneAngleSet = [], None, False
