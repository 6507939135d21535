"""``photonStream`` at the boundary (Code/monteCarloIllumination.f95).

``new_PhotonStream`` keeps the reference's six call forms (:46-50) but returns a DESCRIPTOR: photon
positions and directions are drawn on the device from the batch's Philox streams instead of being
materialised as five host arrays of length N (:87-99).  Assigning the public array components
(xPosition, ..., initialPhi) by hand switches the stream to I3RC_SRC_ARRAYS.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .ErrorMessages import setStateToFailure, setStateToSuccess

_tiny = np.finfo(np.float32).tiny


class photonStream:
    def __init__(self):
        self.currentPhoton = 0
        self.xPosition = self.yPosition = self.zPosition = self.initialMu = self.initialPhi = None
        self.source = None  # _abi.PhotonSource

    def as_c(self) -> _abi.PhotonSource:
        if self.xPosition is not None:
            s = _abi.PhotonSource()
            s.kind = _abi.SRC_ARRAYS
            arrs = [_abi.f32(getattr(self, n)) for n in ("xPosition", "yPosition", "zPosition", "initialMu", "initialPhi")]
            s.numberOfPhotons = arrs[0].size
            s.xPosition, s.yPosition, s.zPosition, s.initialMu, s.initialPhi = [_abi.fptr(a) for a in arrs]
            self._keep = arrs
            return s
        return self.source


def new_PhotonStream(solarMu=None, solarAzimuth=None, solarX=None, solarY=None, *, numberOfPhotons,
                     randomNumbers=None, status=None, detectorX=None, detectorY=None, detectorZ=None,
                     detectorPointsUp=None, detectorMu=None, detectorPhi=None, deltaX=None, deltaY=None):
    """Generic interface of monteCarloIllumination.f95:46-50, resolved on which arguments are present:

    (solarMu, solarAzimuth)                  -> newPhotonStream_Directional   (:62)
    (solarMu)                                -> newPhotonStream_RandomAzimuth (:106)
    ()                                       -> newPhotonStream_Flux          (:148)
    (solarMu, solarAzimuth, solarX, solarY)  -> newPhotonStream_Spotlight     (:187)
    (detectorX,Y,Z, detectorPointsUp)        -> newPhotonStream_Internal_Flux (:228)
    (detectorX,Y,Z, detectorMu, detectorPhi) -> newPhotonStream_Internal_Intensity (:329)
    """
    ph = photonStream()
    s = _abi.PhotonSource()
    s.numberOfPhotons = int(numberOfPhotons)
    if numberOfPhotons <= 0:
        setStateToFailure(status, "setIllumination: must ask for non-negative number of photons.")
        return ph
    if detectorX is not None:
        s.x, s.y, s.z = float(detectorX), float(detectorY), float(detectorZ)
        if detectorMu is not None:
            s.kind, s.detectorMu, s.detectorPhi = _abi.SRC_INTERNAL_INTENSITY, float(detectorMu), float(detectorPhi)
        else:
            s.kind, s.detectorPointsUp = _abi.SRC_INTERNAL_FLUX, int(bool(detectorPointsUp))
        if deltaX is not None:
            s.has_deltaX, s.deltaX = 1, float(deltaX)
        if deltaY is not None:
            s.has_deltaY, s.deltaY = 1, float(deltaY)
    elif solarMu is None:
        s.kind = _abi.SRC_FLUX
    else:
        s.solarMu = float(solarMu)
        if abs(solarMu) > 1.0 or abs(solarMu) <= _tiny:
            setStateToFailure(status, "setIllumination: solarMu out of bounds")
            return ph
        if solarAzimuth is None:
            s.kind = _abi.SRC_RANDOM_AZIMUTH
        else:
            if solarAzimuth < 0.0 or solarAzimuth > 360.0:
                setStateToFailure(status, "setIllumination: solarAzimuth out of bounds")
                return ph
            s.solarAzimuth = float(solarAzimuth)
            if solarX is not None:
                s.kind, s.x, s.y = _abi.SRC_SPOTLIGHT, float(solarX), float(solarY)
            else:
                s.kind = _abi.SRC_DIRECTIONAL
    ph.source, ph.currentPhoton = s, 1
    setStateToSuccess(status)
    return ph


def morePhotonsExist(photons):
    n = photons.source.numberOfPhotons if photons.source is not None else (
        0 if photons.xPosition is None else len(photons.xPosition))
    return 0 < photons.currentPhoton <= n


def finalize_PhotonStream(photons):
    photons.__init__()
