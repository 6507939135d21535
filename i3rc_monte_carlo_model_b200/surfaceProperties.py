"""``surfaceDescription`` at the boundary (Code/surfaceProperties.f95:34-117): a Lambertian albedo map."""
from __future__ import annotations

import numpy as np

from .ErrorMessages import setStateToFailure, setStateToSuccess, stateIsFailure


class surfaceDescription:
    def __init__(self):
        self.xPosition = self.yPosition = self.BRDFParameters = None


def new_SurfaceDescription(surfaceParameters, xPosition=None, yPosition=None, status=None):
    """newSurfaceDescriptionXY (:60) when positions are given, newSurfaceUniform (:98) otherwise."""
    s = surfaceDescription()
    p = np.asarray(surfaceParameters, dtype=np.float32)
    if xPosition is None:
        if p.size != 1:
            setStateToFailure(status, "new_SurfaceDescription: Wrong number of parameters supplied for surface BRDF.")
            return s
        huge = np.finfo(np.float32).max
        xPosition, yPosition, p = [0.0, huge], [0.0, huge], p.reshape(1, 1, 1)
    x = np.asarray(xPosition, dtype=np.float32)
    y = np.asarray(yPosition, dtype=np.float32)
    if p.ndim != 3 or p.shape[0] != 1:
        setStateToFailure(status, "new_SurfaceDescription: Wrong number of parameters supplied for surface BRDF.")
    elif p.shape[1] != x.size - 1 or p.shape[2] != y.size - 1:
        setStateToFailure(status, "new_SurfaceDescription: position vector(s) are incorrect length.")
    if np.any(np.diff(x) <= 0) or np.any(np.diff(y) <= 0):
        setStateToFailure(status, "new_SurfaceDescription: positions must be unique, increasing.")
    if np.any(p < 0) or np.any(p > 1):
        setStateToFailure(status, "new_SurfaceDescription: surface reflectance must be between 0 and 1")
    if status is not None and stateIsFailure(status):
        return s
    s.xPosition, s.yPosition, s.BRDFParameters = x, y, p
    setStateToSuccess(status)
    return s


def isReady_surfaceDescription(s):
    return s.xPosition is not None and s.yPosition is not None and s.BRDFParameters is not None


def finalize_surfaceDescription(s):
    s.xPosition = s.yPosition = s.BRDFParameters = None
