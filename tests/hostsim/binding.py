"""ctypes driver of tests/hostsim/hostsim.cpp (the product's transport core compiled for the CPU).

TEST INFRASTRUCTURE ONLY.  Lets the CPU test-suite check the kernel's logic and Philox streams against the
oracle without a GPU.  Tables (inverse / forward) are taken from the oracle so that only transport differs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from i3rc_monte_carlo_model_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libhostsim.so")
_SRCS = [os.path.join(_HERE, "hostsim.cpp")] + [
    os.path.join(_HERE, "..", "..", "i3rc_monte_carlo_model_b200", "csrc", f) for f in ("transport.cuh", "philox.cuh")]
CNT_NAMES = _abi.COUNTER_FIELDS

fp, ip = _abi.c_float_p, _abi.c_int32_p


class Args(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nc", C.c_int),
        ("xe", fp), ("ye", fp), ("ze", fp), ("ext", fp), ("cum", fp), ("ssa", fp), ("pf", ip),
        ("xyRegular", C.c_int), ("zRegular", C.c_int),
        ("inv", C.POINTER(fp)), ("fwd", C.POINTER(fp)), ("fwdOrig", C.POINTER(fp)),
        ("nInv", ip), ("nFwd", ip), ("nEntries", ip),
        ("nDir", C.c_int), ("mus", fp), ("phisDeg", fp),
        ("useRayTracing", C.c_int), ("useRussianRoulette", C.c_int), ("useRRIntensity", C.c_int),
        ("useHybrid", C.c_int), ("numOrdersOrig", C.c_int), ("limitContrib", C.c_int), ("trackByComponent", C.c_int),
        ("surfaceAlbedo", C.c_float), ("zetaMin", C.c_float), ("maxContrib", C.c_float),
        ("surf_nx", C.c_int), ("surf_ny", C.c_int), ("surf_x", fp), ("surf_y", fp), ("surf_p", fp),
        ("kind", C.c_int), ("n", C.c_longlong),
        ("solarMu", C.c_float), ("solarAzimuthDeg", C.c_float), ("sx", C.c_float), ("sy", C.c_float), ("sz", C.c_float),
        ("detectorMu", C.c_float), ("detectorPhi", C.c_float),
        ("pointsUp", C.c_int), ("hasDx", C.c_int), ("hasDy", C.c_int), ("deltaX", C.c_float), ("deltaY", C.c_float),
        ("ax", fp), ("ay", fp), ("az", fp), ("amu", fp), ("aphi", fp),
        ("key0", C.c_uint32), ("key1", C.c_uint32),
        ("fluxUp", fp), ("fluxDown", fp), ("fluxAbs", fp), ("volAbs", fp), ("intensity", fp), ("intByComp", fp),
        ("excess", fp), ("counters", C.POINTER(C.c_ulonglong)),
    ]


def lib():
    if not os.path.exists(_LIB) or any(os.path.getmtime(_LIB) < os.path.getmtime(s) for s in _SRCS):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _LIB, _SRCS[0]])
    L = C.CDLL(_LIB)
    L.hostsim_run.argtypes = [C.POINTER(Args)]
    L.hostsim_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    L.hostsim_trace_rays.argtypes = [C.POINTER(Args), C.c_int, fp, fp, fp, fp, fp, ip]
    L.hostsim_trace_rays_jump.argtypes = [C.POINTER(Args), C.c_int, fp, fp, fp, fp, fp, ip, C.POINTER(C.c_ulonglong)]
    L.hostsim_trace_rays_slab.argtypes = [C.POINTER(Args), C.c_int, C.c_int, fp, fp, fp, fp, fp, ip, C.POINTER(C.c_ulonglong)]
    L.hostsim_set_jump.argtypes = [C.c_int]
    L.hostsim_set_vertical.argtypes = [C.c_int]
    L.hostsim_set_lower_bound.argtypes = [C.c_int]
    return L


def philox(counter, key):
    out = (C.c_uint32 * 4)()
    lib().hostsim_philox(*[int(c) for c in counter], int(key[0]), int(key[1]), out)
    return [int(v) for v in out]


def dense_from_domain(d):
    """getOpticalPropertiesByComponent (Code/opticalProperties.f95:429-539) in numpy float32; [x,y,z(,c)] arrays."""
    nx, ny, nz = d.xPosition.size - 1, d.yPosition.size - 1, d.zPosition.size - 1
    nc = len(d.components)
    cum = np.zeros((nx, ny, nz, nc), np.float32, order="F")
    ssa = np.zeros_like(cum)
    pfi = np.zeros((nx, ny, nz, nc), np.int32, order="F")
    for c, k in enumerate(d.components):
        z0, z1 = k.zLevelBase - 1, k.zLevelBase - 1 + k.extinction.shape[2]
        cum[:, :, z0:z1, c] = np.broadcast_to(k.extinction, (nx, ny, z1 - z0))
        ssa[:, :, z0:z1, c] = np.broadcast_to(k.singleScatteringAlbedo, (nx, ny, z1 - z0))
        pfi[:, :, z0:z1, c] = np.broadcast_to(k.phaseFunctionIndex, (nx, ny, z1 - z0))
    for c in range(1, nc):
        cum[..., c] = cum[..., c] + cum[..., c - 1]
    tot = np.asfortranarray(cum[..., nc - 1].copy())
    m = tot > np.finfo(np.float32).tiny
    for c in range(nc):
        cum[..., c][m] = cum[..., c][m] / tot[m]
    eps = np.float32(np.finfo(np.float32).eps)
    last = cum[..., nc - 1]
    last[np.abs(last - 1) <= eps] = np.float32(1) + eps
    return tot, cum, ssa, pfi


class HostSim:
    """Runs batches of the product's transport core on the CPU for one problem."""

    def __init__(self, domain, oracle_integrator, getTable):
        self.L = lib()
        d = domain
        self.nx, self.ny, self.nz = d.xPosition.size - 1, d.yPosition.size - 1, d.zPosition.size - 1
        self.nc = len(d.components)
        self.xe, self.ye, self.ze = (_abi.f32(a) for a in (d.xPosition, d.yPosition, d.zPosition))
        self.tot, self.cum, self.ssa, self.pfi = dense_from_domain(d)
        self.xyReg, self.zReg = int(d.xyRegularlySpaced), int(d.zRegularlySpaced)
        self.inv = [getTable(oracle_integrator, 0, c) for c in range(self.nc)]
        try:
            self.fwd = [getTable(oracle_integrator, 1, c) for c in range(self.nc)]
            self.fwdO = [getTable(oracle_integrator, 2, c) for c in range(self.nc)]
        except RuntimeError:
            self.fwd = self.fwdO = None

    def _args(self, photons, seed, **kw):
        a = Args()
        keep = []
        a.nx, a.ny, a.nz, a.nc = self.nx, self.ny, self.nz, self.nc
        a.xe, a.ye, a.ze = _abi.fptr(self.xe), _abi.fptr(self.ye), _abi.fptr(self.ze)
        a.ext, a.cum, a.ssa, a.pf = _abi.fptr(self.tot), _abi.fptr(self.cum), _abi.fptr(self.ssa), _abi.iptr(self.pfi)
        a.xyRegular, a.zRegular = self.xyReg, self.zReg
        nc = self.nc
        inv = (fp * nc)(*[_abi.fptr(t) for t in self.inv])
        a.inv = inv
        nInv = _abi.i32([t.shape[1] for t in self.inv])
        nEnt = _abi.i32([t.shape[0] for t in self.inv])
        a.nInv, a.nEntries = _abi.iptr(nInv), _abi.iptr(nEnt)
        keep += [inv, nInv, nEnt]
        mus = kw.get("intensityMus")
        nD = 0
        if mus is not None:
            mus, phis = _abi.f32(mus), _abi.f32(kw["intensityPhis"])
            nD = mus.size
            fwd = (fp * nc)(*[_abi.fptr(t) for t in self.fwd])
            fwdO = (fp * nc)(*[_abi.fptr(t) for t in self.fwdO])
            nFwd = _abi.i32([t.shape[1] for t in self.fwd])
            a.fwd, a.fwdOrig, a.nFwd, a.mus, a.phisDeg = fwd, fwdO, _abi.iptr(nFwd), _abi.fptr(mus), _abi.fptr(phis)
            keep += [mus, phis, fwd, fwdO, nFwd]
        a.nDir = nD
        a.useRayTracing = int(kw.get("useRayTracing", True))
        a.useRussianRoulette = int(kw.get("useRussianRoulette", True))
        a.useRRIntensity = int(kw.get("useRussianRouletteForIntensity", False))
        a.useHybrid = int(kw.get("useHybridPhaseFunsForIntenCalcs", False))
        a.numOrdersOrig = int(kw.get("numOrdersOrigPhaseFunIntenCalcs", 0))
        a.limitContrib = int(kw.get("limitIntensityContributions", False))
        a.trackByComponent = int(kw.get("trackByComponent", False))
        a.surfaceAlbedo = float(kw.get("surfaceAlbedo", 0.0))
        a.zetaMin = float(kw.get("zetaMin", 0.3))
        a.maxContrib = float(kw.get("maxIntensityContribution", np.finfo(np.float32).max))
        s = photons.as_c()
        a.kind, a.n = s.kind, s.numberOfPhotons
        a.solarMu, a.solarAzimuthDeg, a.sx, a.sy, a.sz = s.solarMu, s.solarAzimuth, s.x, s.y, s.z
        a.detectorMu, a.detectorPhi, a.pointsUp = s.detectorMu, s.detectorPhi, s.detectorPointsUp
        a.hasDx, a.hasDy, a.deltaX, a.deltaY = s.has_deltaX, s.has_deltaY, s.deltaX, s.deltaY
        a.ax, a.ay, a.az, a.amu, a.aphi = s.xPosition, s.yPosition, s.zPosition, s.initialMu, s.initialPhi
        keep.append(photons)
        a.key0, a.key1 = int(seed[0]) & 0xFFFFFFFF, int(seed[1]) & 0xFFFFFFFF
        ncol, ncell = self.nx * self.ny, self.nx * self.ny * self.nz
        out = {
            "fluxUp": np.zeros(ncol, np.float32), "fluxDown": np.zeros(ncol, np.float32),
            "fluxAbsorbed": np.zeros(ncol, np.float32), "volumeAbsorption": np.zeros(ncell, np.float32),
            "intensity": np.zeros(max(ncol * nD, 1), np.float32),
            "intByComp": np.zeros(max(ncol * nD * (nc + 1), 1), np.float32),
            "excess": np.zeros(max(nD * (nc + 1), 1), np.float32),
        }
        a.fluxUp, a.fluxDown, a.fluxAbs, a.volAbs = (_abi.fptr(out[k]) for k in ("fluxUp", "fluxDown", "fluxAbsorbed", "volumeAbsorption"))
        a.intensity, a.intByComp, a.excess = _abi.fptr(out["intensity"]), _abi.fptr(out["intByComp"]), _abi.fptr(out["excess"])
        cnt = (C.c_ulonglong * len(CNT_NAMES))()
        a.counters = cnt
        return a, out, cnt, keep, nD

    def run(self, photons, seed, **kw):
        """One batch; returns normalised results shaped like reportResults ([x,y(,..)] Fortran order) + counters."""
        a, out, cnt, keep, nD = self._args(photons, seed, **kw)
        self.L.hostsim_run(C.byref(a))
        n = float(np.float32(a.n))
        nx, ny, nz = self.nx, self.ny, self.nz
        nppc = np.float32(n) / np.float32(nx * ny)
        dz = np.diff(self.ze)
        res = {
            "fluxUp": (out["fluxUp"] / nppc).reshape(ny, nx).T, "fluxDown": (out["fluxDown"] / nppc).reshape(ny, nx).T,
            "fluxAbsorbed": (out["fluxAbsorbed"] / nppc).reshape(ny, nx).T,
            "volumeAbsorption": (out["volumeAbsorption"].reshape(nz, ny, nx) / (nppc * dz[:, None, None])).transpose(2, 1, 0),
        }
        if nD:
            res["intensity"] = (out["intensity"][: nx * ny * nD] / nppc).reshape(nD, ny, nx).transpose(2, 1, 0)
            res["meanIntensity"] = res["intensity"].mean(axis=(0, 1))
        for k in ("fluxUp", "fluxDown", "fluxAbsorbed"):
            res["mean" + k[0].upper() + k[1:]] = float(res[k].mean())
        res["absorbedProfile"] = res["volumeAbsorption"].mean(axis=(0, 1))
        res["counters"] = {n_: int(cnt[i]) for i, n_ in enumerate(CNT_NAMES)}
        return res

    def trace_rays_slab(self, pos, direction, tauLimit=None, jump=True):
        """The same through the library's layer-compacted field (regular grids with horizontally uniform layers): runs of
        uniform layers crossed in one go (jump) or cell by cell.  Returns tau, end points, end cells, (skipped, steps)."""
        from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
        a, out, cnt, keep, nD = self._args(new_PhotonStream(0.5, 0.0, numberOfPhotons=1), (0, 0))
        pos, direction = _abi.f32(pos).reshape(-1, 3), _abi.f32(direction).reshape(-1, 3)
        n = pos.shape[0]
        tau, pout, idx = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.int32)
        lim = _abi.f32(tauLimit) if tauLimit is not None else None
        sk = (C.c_ulonglong * 2)()
        rc = self.L.hostsim_trace_rays_slab(C.byref(a), int(jump), n, _abi.fptr(pos), _abi.fptr(direction), _abi.fptr(lim),
                                            _abi.fptr(tau), _abi.fptr(pout), _abi.iptr(idx), sk)
        assert rc == 0, "needs a regular grid with horizontally uniform layers"
        return tau, pout, idx, (int(sk[0]), int(sk[1]))

    def trace_rays(self, pos, direction, tauLimit=None, jump=False):
        """accumulateExtinctionAlongPath for explicit rays; jump=True: through the field with empty-space codes (regular
        grids), then also returns (cells skipped, DDA steps taken)."""
        from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
        a, out, cnt, keep, nD = self._args(new_PhotonStream(0.5, 0.0, numberOfPhotons=1), (0, 0))
        pos, direction = _abi.f32(pos).reshape(-1, 3), _abi.f32(direction).reshape(-1, 3)
        n = pos.shape[0]
        tau, pout, idx = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.int32)
        lim = _abi.f32(tauLimit) if tauLimit is not None else None
        if jump:
            sk = (C.c_ulonglong * 2)()
            rc = self.L.hostsim_trace_rays_jump(C.byref(a), n, _abi.fptr(pos), _abi.fptr(direction), _abi.fptr(lim), _abi.fptr(tau),
                                                _abi.fptr(pout), _abi.iptr(idx), sk)
            assert rc == 0, "empty-space codes need a regular grid"
            return tau, pout, idx, (int(sk[0]), int(sk[1]))
        self.L.hostsim_trace_rays(C.byref(a), n, _abi.fptr(pos), _abi.fptr(direction), _abi.fptr(lim), _abi.fptr(tau),
                                  _abi.fptr(pout), _abi.iptr(idx))
        return tau, pout, idx

    def set_lower_bound(self, on):
        """Local-estimate rays whose lower bound of the optical path to the top exceeds their roulette budget are not traced."""
        self.L.hostsim_set_lower_bound(int(on))

    def set_vertical(self, on):
        """Radiance directions that point straight up are integrated from column suffix sums (Problem::colTau)."""
        self.L.hostsim_set_vertical(int(on))

    def set_jump(self, on):
        """Photon batches (run) use the empty-space codes too (regular grids, ray tracing)."""
        self.L.hostsim_set_jump(int(on))
