"""``domain`` at the boundary (Code/opticalProperties.f95:54-357).

The host object only validates and holds the component arrays; the expansion that
``getOpticalPropertiesByComponent`` (:429-539) performs -- dense totalExt / cumulativeExt / ssa /
phaseFunctionIndex -- runs on the device inside ``i3rc_new_Integrator_components``.
Arrays are indexed like the reference, ``a[ix, iy, iz]``.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .ErrorMessages import setStateToFailure, setStateToSuccess, stateIsFailure
from .scatteringPhaseFunctions import isReady_PhaseFunctionTable


class opticalComponent:
    def __init__(self, name, extinction, ssa, phaseFunctionIndex, zLevelBase, table):
        self.name, self.zLevelBase, self.table = name, int(zLevelBase), table
        self.horizontallyUniform = extinction.shape[0] == 1 and extinction.shape[1] == 1
        self.extinction = np.asfortranarray(extinction, dtype=np.float32)
        self.singleScatteringAlbedo = np.asfortranarray(ssa, dtype=np.float32)
        self.phaseFunctionIndex = np.asfortranarray(phaseFunctionIndex, dtype=np.int32)


class domain:
    def __init__(self):
        self.xPosition = self.yPosition = self.zPosition = None
        self.xyRegularlySpaced = self.zRegularlySpaced = False
        self.components: list[opticalComponent] = []


def _spacing(a):
    return np.spacing(np.abs(np.asarray(a, np.float32)))


def new_Domain(xPosition, yPosition, zPosition, status=None):
    """opticalProperties.f95:93-131"""
    d = domain()
    x, y, z = (np.asarray(a, dtype=np.float32).ravel() for a in (xPosition, yPosition, zPosition))
    if np.any(np.diff(x) <= 0) or np.any(np.diff(y) <= 0) or np.any(np.diff(z) <= 0):
        setStateToFailure(status, "new_Domain: Positions must be increasing, unique.")
        return d
    d.xPosition, d.yPosition, d.zPosition = x.copy(), y.copy(), z.copy()
    d.xyRegularlySpaced = bool(np.all(np.abs(np.diff(x) - (x[1] - x[0])) <= 2 * _spacing(x[1:])) and
                               np.all(np.abs(np.diff(y) - (y[1] - y[0])) <= 2 * _spacing(y[1:])))
    d.zRegularlySpaced = bool(np.all(np.abs(np.diff(z) - (z[1] - z[0])) <= 2 * _spacing(z[1:])))
    setStateToSuccess(status)
    return d


def _isValid(d):
    return d.xPosition is not None


def _validateOpticalComponent(d, extinction, ssa, pfi, table, zLevelBase, status):
    """opticalProperties.f95:929-987"""
    if not _isValid(d):
        setStateToFailure(status, "validateOpticalComponent: domain hasn't been initialized.")
        return
    nX, nY, nZ = extinction.shape
    if ssa.shape != extinction.shape or pfi.shape != extinction.shape:
        setStateToFailure(status, "validateOpticalComponent: optical property grids must be the same size.")
    if nX not in (1, d.xPosition.size - 1) or nY not in (1, d.yPosition.size - 1):
        setStateToFailure(status, "validateOpticalComponent: arrays don't conform to horizontal extent of domain.")
    if zLevelBase + nZ - 1 > d.zPosition.size or zLevelBase < 1:
        setStateToFailure(status, "validateOpticalComponent: arrays don't conform to vertical extent of domain.")
    if np.any(extinction < 0):
        setStateToFailure(status, "validateOpticalComponent: extinction must be >= 0.")
    if np.any(ssa < 0) or np.any(ssa > 1):
        setStateToFailure(status, "validateOpticalComponent: singleScatteringAlbedo must be between 0 and 1")
    if status is None or not stateIsFailure(status):
        n = len(table.phaseFunctions)
        if np.any(pfi < 0) or np.any(pfi > n):
            setStateToFailure(status, "validateOpticalComponent: phase function index is out of bounds")
        if not isReady_PhaseFunctionTable(table):
            setStateToFailure(status, "validateOpticalComponent: phase function table is not ready.")


def _as3d(a, dtype):
    a = np.asarray(a, dtype=dtype)
    return a.reshape(1, 1, -1) if a.ndim == 1 else a  # addOpticalComponent1D, :198-230


def addOpticalComponent(thisDomain, componentName, extinction, singleScatteringAlbedo, phaseFunctionIndex,
                        phaseFunctions, zLevelBase=1, status=None):
    """opticalProperties.f95:133-230 (3D and 1D forms)."""
    e, s, p = _as3d(extinction, np.float32), _as3d(singleScatteringAlbedo, np.float32), _as3d(phaseFunctionIndex, np.int32)
    _validateOpticalComponent(thisDomain, e, s, p, phaseFunctions, zLevelBase, status)
    if status is not None and stateIsFailure(status):
        setStateToFailure(status, "addOpticalComponent: optical properties aren't valid.")
        return
    thisDomain.components.append(opticalComponent(componentName, e, s, p, zLevelBase, phaseFunctions))
    setStateToSuccess(status)


def replaceOpticalComponent(thisDomain, componentNumber, componentName, extinction, singleScatteringAlbedo,
                            phaseFunctionIndex, phaseFunctions, zLevelBase=1, status=None):
    """opticalProperties.f95:232-309; componentNumber is 1-based."""
    if not thisDomain.components or not (1 <= componentNumber <= len(thisDomain.components)):
        setStateToFailure(status, "replaceOpticalComponent: no components to replace")
        return
    e, s, p = _as3d(extinction, np.float32), _as3d(singleScatteringAlbedo, np.float32), _as3d(phaseFunctionIndex, np.int32)
    _validateOpticalComponent(thisDomain, e, s, p, phaseFunctions, zLevelBase, status)
    if status is not None and stateIsFailure(status):
        return
    thisDomain.components[componentNumber - 1] = opticalComponent(componentName, e, s, p, zLevelBase, phaseFunctions)
    setStateToSuccess(status)


def deleteOpticalComponent(thisDomain, componentNumber, status=None):
    """opticalProperties.f95:311-357"""
    if not _isValid(thisDomain):
        setStateToFailure(status, "deleteOpticalComponent: domain hasn't been initialized.")
    elif not thisDomain.components:
        setStateToFailure(status, "deleteOpticalComponent: no components to delete.")
    elif not (1 <= componentNumber <= len(thisDomain.components)):
        setStateToFailure(status, "deleteOpticalComponent: non-existent component.")
    else:
        del thisDomain.components[componentNumber - 1]
        setStateToSuccess(status)


def getInfo_Domain(thisDomain, status=None):
    """opticalProperties.f95:361-425"""
    if not _isValid(thisDomain):
        setStateToFailure(status, "getInfo_Domain: domain hasn't been initialized.")
        return {}
    setStateToSuccess(status)
    return {
        "numX": thisDomain.xPosition.size - 1, "numY": thisDomain.yPosition.size - 1,
        "numZ": thisDomain.zPosition.size - 1,
        "xPosition": thisDomain.xPosition, "yPosition": thisDomain.yPosition, "zPosition": thisDomain.zPosition,
        "numberOfComponents": len(thisDomain.components),
        "componentNames": [c.name for c in thisDomain.components],
    }


def finalize_Domain(thisDomain):
    thisDomain.__init__()


def components_as_c(thisDomain):
    """Array of i3rc_component for i3rc_new_Integrator_components; returns (array, keepalive)."""
    n = len(thisDomain.components)
    arr = (_abi.Component * n)()
    keep = []
    for i, c in enumerate(thisDomain.components):
        arr[i].extinction = _abi.fptr(c.extinction)
        arr[i].ssa = _abi.fptr(c.singleScatteringAlbedo)
        arr[i].phase_index = _abi.iptr(c.phaseFunctionIndex)
        arr[i].horizontally_uniform = int(c.horizontallyUniform)
        arr[i].z_level_base = c.zLevelBase
        arr[i].nz = c.extinction.shape[2]
        arr[i].table = c.table.as_c()
        keep.append((c, c.table._keep))
    return arr, keep
