"""Host-side mirror of module monteCarloRadiativeTransfer (Integrators/monteCarloRadiativeTransfer.f95).

Same public list as the reference (MCRT:154-156): ``integrator``, ``new_Integrator``, ``copy_Integrator``,
``isReady_Integrator``, ``finalize_Integrator``, ``specifyParameters``, ``computeRadiativeTransfer``,
``reportResults``.  Every call goes straight through the C ABI of include/i3rc_b200.h; argument
validation and the status text live in the library, not here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .ErrorMessages import from_return_code, setStateToCompleteSuccess, setStateToFailure
from .opticalProperties import components_as_c


class integrator:
    def __init__(self, backend, handle):
        self.backend, self.handle = backend, handle
        self.nx = self.ny = self.nz = self.nc = 0
        self.nDir = 0

    def _msg(self):
        m = self.backend.last_message(self.handle)
        return m.decode() if m else ""

    def __del__(self):
        try:
            finalize_Integrator(self)
        except Exception:
            pass


def _default_backend():
    from ._lib import backend
    return backend()


def new_Integrator(atmosphere, status=None, backend=None):
    """MCRT:162-254.  ``atmosphere`` is a ``domain``; the dense property arrays are built on the device."""
    be = backend or _default_backend()
    h = C.c_void_p()
    if atmosphere.xPosition is None or not atmosphere.components:
        setStateToFailure(status, "new_Integrator: Problems reading domain.")
        return integrator(be, None)
    x, y, z = (_abi.f32(a) for a in (atmosphere.xPosition, atmosphere.yPosition, atmosphere.zPosition))
    comps, keep = components_as_c(atmosphere)
    rc = be.new_Integrator_components(x.size - 1, y.size - 1, z.size - 1, _abi.fptr(x), _abi.fptr(y), _abi.fptr(z),
                                      len(comps), comps, C.byref(h))
    new = integrator(be, h if rc != _abi.FAILURE else None)
    new.nx, new.ny, new.nz, new.nc = x.size - 1, y.size - 1, z.size - 1, len(comps)
    from_return_code(status, rc, new._msg() if new.handle else "new_Integrator: Problems reading domain.")
    return new


def new_Integrator_dense(xPosition, yPosition, zPosition, totalExt, cumulativeExt, ssa, phaseFunctionIndex,
                         tables, status=None, backend=None):
    """The dense form of new_Integrator: arrays as getOpticalPropertiesByComponent returns them
    (Code/opticalProperties.f95:429-539), indexed [ix, iy, iz(, component)]."""
    be = backend or _default_backend()
    x, y, z = (_abi.f32(a) for a in (xPosition, yPosition, zPosition))
    te = np.asfortranarray(totalExt, np.float32)
    ce = np.asfortranarray(cumulativeExt, np.float32)
    sa = np.asfortranarray(ssa, np.float32)
    pi = np.asfortranarray(phaseFunctionIndex, np.int32)
    nc = ce.shape[3]
    h = C.c_void_p()
    rc = be.new_Integrator(x.size - 1, y.size - 1, z.size - 1, nc, _abi.fptr(x), _abi.fptr(y), _abi.fptr(z),
                           _abi.fptr(te), _abi.fptr(ce), _abi.fptr(sa), _abi.iptr(pi), C.byref(h))
    new = integrator(be, h if rc != _abi.FAILURE else None)
    new.nx, new.ny, new.nz, new.nc = x.size - 1, y.size - 1, z.size - 1, nc
    if new.handle:
        for c, t in enumerate(tables):
            ct = t.as_c()
            rc = max(rc, be.set_phase_table(new.handle, c, C.byref(ct)))
    from_return_code(status, rc, new._msg() if new.handle else "new_Integrator: Problems reading domain.")
    return new


def copy_Integrator(original):
    """MCRT:1082-1253"""
    h = C.c_void_p()
    original.backend.copy_Integrator(original.handle, C.byref(h))
    cp = integrator(original.backend, h)
    cp.nx, cp.ny, cp.nz, cp.nc, cp.nDir = original.nx, original.ny, original.nz, original.nc, original.nDir
    return cp


def isReady_Integrator(thisIntegrator):
    return bool(thisIntegrator.handle) and bool(thisIntegrator.backend.isReady_Integrator(thisIntegrator.handle))


def finalize_Integrator(thisIntegrator):
    """MCRT:1258-1349"""
    if getattr(thisIntegrator, "handle", None):
        thisIntegrator.backend.finalize_Integrator(thisIntegrator.handle)
        thisIntegrator.handle = None


_SCALARS = ("surfaceAlbedo", "minForwardTableSize", "minInverseTableSize", "computeIntensity", "useRayTracing",
            "useRussianRoulette", "useRussianRouletteForIntensity", "zetaMin", "useHybridPhaseFunsForIntenCalcs",
            "hybridPhaseFunWidth", "numOrdersOrigPhaseFunIntenCalcs", "limitIntensityContributions",
            "maxIntensityContribution")


def specifyParameters(thisIntegrator, surfaceAlbedo=None, surfaceBDRF=None, minForwardTableSize=None,
                      minInverseTableSize=None, intensityMus=None, intensityPhis=None, computeIntensity=None,
                      useRayTracing=None, useRussianRoulette=None, useRussianRouletteForIntensity=None, zetaMin=None,
                      useHybridPhaseFunsForIntenCalcs=None, hybridPhaseFunWidth=None,
                      numOrdersOrigPhaseFunIntenCalcs=None, limitIntensityContributions=None,
                      maxIntensityContribution=None, status=None):
    """MCRT:830-1069; ``None`` = optional argument not present."""
    p = _abi.Params()
    loc = locals()
    keep = []
    for name in _SCALARS:
        v = loc[name]
        if v is not None:
            p.present |= _abi.P_BITS[name]
            setattr(p, name, v)
    if intensityMus is not None:
        mus = _abi.f32(np.atleast_1d(intensityMus))
        p.present |= _abi.P_BITS["intensityMus"]
        p.intensityMus, p.numIntensityDirections = _abi.fptr(mus), mus.size
        keep.append(mus)
    if intensityPhis is not None:
        phis = _abi.f32(np.atleast_1d(intensityPhis))
        p.present |= _abi.P_BITS["intensityPhis"]
        p.intensityPhis = _abi.fptr(phis)
        if intensityMus is None:
            p.numIntensityDirections = phis.size
        elif phis.size != p.numIntensityDirections:
            setStateToFailure(status, "specifyParameters: intensityMus, intensityPhis must be the same length.")
            return
        keep.append(phis)
    if surfaceBDRF is not None:
        p.present |= _abi.P_BITS["surfaceBDRF"]
        if surfaceBDRF.xPosition is not None:
            sx, sy = _abi.f32(surfaceBDRF.xPosition), _abi.f32(surfaceBDRF.yPosition)
            sp = np.asfortranarray(surfaceBDRF.BRDFParameters[0], np.float32)
            p.surf_nx, p.surf_ny = sx.size - 1, sy.size - 1
            p.surf_x, p.surf_y, p.surf_params = _abi.fptr(sx), _abi.fptr(sy), _abi.fptr(sp)
            keep += [sx, sy, sp]
    rc = thisIntegrator.backend.specifyParameters(thisIntegrator.handle, C.byref(p))
    if rc != _abi.FAILURE and intensityMus is not None:
        thisIntegrator.nDir = p.numIntensityDirections
    if rc != _abi.FAILURE and computeIntensity is False and intensityMus is None:
        thisIntegrator.nDir = 0
    from_return_code(status, rc, thisIntegrator._msg())


def computeRadiativeTransfer(thisIntegrator, randomNumbers, incomingPhotons, status=None):
    """MCRT:262-398.  ``randomNumbers`` carries the seed vector; ``incomingPhotons`` the source descriptor."""
    if not isReady_Integrator(thisIntegrator):
        setStateToFailure(status, "computeRadiativeTransfer: problem not completely specified.")
        return
    src = incomingPhotons.as_c()
    if src is None:
        setStateToFailure(status, "getNextPhoton: photons have not been initialized.")
        return
    seed = _abi.i32(randomNumbers.seed)
    rc = thisIntegrator.backend.computeRadiativeTransfer(thisIntegrator.handle, C.byref(src), _abi.iptr(seed), seed.size)
    if rc == _abi.SUCCESS:
        setStateToCompleteSuccess(status, "computeRadiativeTransfer: finished with photons")
        incomingPhotons.currentPhoton = int(src.numberOfPhotons) + 1
    else:
        from_return_code(status, rc, thisIntegrator._msg())


_REPORT = ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxUp", "fluxDown", "fluxAbsorbed", "absorbedProfile",
           "volumeAbsorption", "meanIntensity", "intensity")


def reportResults(thisIntegrator, *want, status=None, out=None):
    """MCRT:711-826.  ``want`` names the optional arguments that are present; returns them in a dict.

    ``out`` may supply caller-allocated float32 Fortran-ordered arrays (size-checked like MCRT:746-791).
    """
    I = thisIntegrator
    want = want or _REPORT[:7]
    shapes = {
        "meanFluxUp": (1,), "meanFluxDown": (1,), "meanFluxAbsorbed": (1,),
        "fluxUp": (I.nx, I.ny), "fluxDown": (I.nx, I.ny), "fluxAbsorbed": (I.nx, I.ny),
        "absorbedProfile": (I.nz,), "volumeAbsorption": (I.nx, I.ny, I.nz),
        "meanIntensity": (I.nDir,), "intensity": (I.nx, I.ny, I.nDir),
    }
    bufs, args = {}, []
    for name in _REPORT:
        if name in want:
            if out is not None and name in out:
                a = out[name]
                if tuple(a.shape) != shapes[name]:
                    setStateToFailure(status, f"reportResults: {name} array is the wrong size")
                    return {}
            else:
                a = np.zeros(shapes[name], dtype=np.float32, order="F")
            bufs[name] = a
            args.append(_abi.fptr(a))
        else:
            args.append(None)
    rc = I.backend.reportResults(I.handle, *args)
    from_return_code(status, rc, I._msg())
    if rc == _abi.FAILURE:
        return {}
    return {k: (float(v[0]) if k.startswith("meanFlux") else v) for k, v in bufs.items()}


# ---- helpers that are not part of the reference module but of the C ABI -------------------------
def setComponentProfile(thisIntegrator, componentNumber, extinction, status=None):
    """Replace the extinction of component ``componentNumber`` (1-based) by a horizontally uniform profile
    extinction[nZ] on the device -- one k-distribution term of a gas component (i3rc_set_component_profile)."""
    be = thisIntegrator.backend
    if be.set_component_profile is None:
        setStateToFailure(status, "setComponentProfile: not available on this backend.")
        return
    e = _abi.f32(np.asarray(extinction).ravel())
    if e.size != thisIntegrator.nz:
        setStateToFailure(status, "setComponentProfile: extinction must have one value per layer.")
        return
    rc = be.set_component_profile(thisIntegrator.handle, int(componentNumber) - 1, _abi.fptr(e))
    from_return_code(status, rc, thisIntegrator._msg())
    if status is None and rc == _abi.FAILURE:
        raise RuntimeError(thisIntegrator._msg())


def getCounters(thisIntegrator):
    c = _abi.Counters()
    thisIntegrator.backend.get_counters(thisIntegrator.handle, C.byref(c))
    return c.as_dict()


def getTable(thisIntegrator, which, comp=0):
    n, e = C.c_int(), C.c_int()
    be = thisIntegrator.backend
    if be.get_table(thisIntegrator.handle, which, comp, None, C.byref(n), C.byref(e)) == _abi.FAILURE:
        raise RuntimeError("table not available: " + thisIntegrator._msg())
    out = np.zeros((e.value, n.value), np.float32)
    be.get_table(thisIntegrator.handle, which, comp, _abi.fptr(out), C.byref(n), C.byref(e))
    return out


def traceRays(thisIntegrator, pos, direction, tauLimit=None):
    """accumulateExtinctionAlongPath (MCRT:1654-1807) for a set of rays; returns tau, end position, 1-based cell."""
    pos, direction = _abi.f32(pos).reshape(-1, 3), _abi.f32(direction).reshape(-1, 3)
    n = pos.shape[0]
    tau, pout, idx = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.int32)
    lim = _abi.f32(tauLimit) if tauLimit is not None else None
    rc = thisIntegrator.backend.trace_rays(thisIntegrator.handle, n, _abi.fptr(pos), _abi.fptr(direction),
                                           _abi.fptr(lim), _abi.fptr(tau), _abi.fptr(pout), _abi.iptr(idx))
    if rc == _abi.FAILURE:
        raise RuntimeError(thisIntegrator._msg())
    return tau, pout, idx
