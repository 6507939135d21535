"""Domain builders for the benchmark configurations.

Each builder restates, in float32, the recipe of the reference program that writes the domain file:
  C1 planeParallel  Example-Drivers/planeParallel.f95:299-372 (createDomain)
  C2 StepCloud      I3RC-Examples/i3rcStepCloud.f95:54-81
  C3 LandsatCloud   I3RC-Examples/i3rcLandsatCloud.f95:61-123   (data/i3rc_fields.npz)
  C4 RadarCloud     I3RC-Examples/i3rcRadarCloud.f95:66-131     (data/i3rc_fields.npz)
  C5 synthetic LES  no reference recipe: lognormal, Fourier-filtered liquid water in a cloud layer,
                    a multi-entry Henyey-Greenstein table keyed by effective radius, plus a
                    horizontally uniform absorbing gas component (ssa = 0), the reference's only
                    mechanism for gas absorption (Tools/PhysicalPropertiesToDomain.f95:330-347).
"""
from __future__ import annotations

import os

import numpy as np

from .opticalProperties import addOpticalComponent, new_Domain
from .scatteringPhaseFunctions import new_PhaseFunction, new_PhaseFunctionTable

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "i3rc_fields.npz")
f32 = np.float32


def _hg_table(g=0.85, nLegendreCoefficients=64, description="Henyey-Greenstein"):
    coefs = (f32(g) ** np.arange(1, nLegendreCoefficients + 1, dtype=np.int32)).astype(f32)
    phase = new_PhaseFunction(coefs)
    return new_PhaseFunctionTable([phase], [1.0], tableDescription=description)


def plane_parallel(nX=1, nY=1, nLayers=1, domainSize=500.0, physicalThickness=250.0, opticalDepth=1.0, SSA=1.0,
                   g=0.85, nLegendreCoefficients=64, useMoments=True, nAngles=5000):
    x = f32(domainSize) / f32(nX) * np.arange(0, nX + 1, dtype=f32)
    y = f32(domainSize) / f32(nY) * np.arange(0, nY + 1, dtype=f32)
    z = f32(physicalThickness) / f32(nLayers) * np.arange(0, nLayers + 1, dtype=f32)
    d = new_Domain(x, y, z)
    if useMoments:
        table = _hg_table(g, nLegendreCoefficients)
    else:
        ang = (np.arange(nAngles, dtype=f32) / f32(nAngles - 1) * np.arccos(f32(-1.0))).astype(f32)
        val = ((1 - f32(g) ** 2) / (1 + f32(g) ** 2 - 2 * f32(g) * np.cos(ang)) ** f32(1.5)).astype(f32)
        table = new_PhaseFunctionTable(ang, val[:, None], [1.0])
    ext = np.full((nX, nY, nLayers), f32(opticalDepth) / f32(physicalThickness), f32)
    addOpticalComponent(d, "cloud", ext, np.full_like(ext, SSA), np.ones(ext.shape, np.int32), table)
    return d


def step_cloud(SSA=1.0, nLegendreCoefficients=64):
    nColumns, nLayers, domainSize, thick = 32, 32, f32(500.0), f32(250.0)
    deltaX, deltaZ = domainSize / f32(nColumns), thick / f32(nLayers)
    d = new_Domain(deltaX * np.arange(0, nColumns + 1, dtype=f32), np.array([0.0, 500.0], f32),
                   deltaZ * np.arange(0, nLayers + 1, dtype=f32))
    tau = np.array([2] * (nColumns // 2) + [18] * (nColumns // 2), f32)
    ext = np.repeat((tau / thick)[:, None, None], nLayers, axis=2).astype(f32)
    addOpticalComponent(d, "cloud", ext, np.full_like(ext, SSA), np.ones(ext.shape, np.int32),
                        _hg_table(0.85, nLegendreCoefficients))
    return d


def landsat_cloud(SSA=1.0, nLegendreCoefficients=299):
    dat = np.load(_DATA)
    nX = nY = 128
    deltaXY, deltaZ, maxThickness = f32(30.0), 20, 2380
    nLayers = (maxThickness + deltaZ // 2) // deltaZ  # 119
    opticalDepth = dat["landsat_tau"].T.astype(f32)    # file rows are y: opticalDepth(:, i) per row
    thickness = (dat["landsat_dz_km"].T * f32(1000.0)).astype(f32)
    d = new_Domain(deltaXY * np.arange(0, nX + 1, dtype=f32), deltaXY * np.arange(0, nY + 1, dtype=f32),
                   f32(deltaZ) * np.arange(0, nLayers + 1, dtype=f32) + f32(200))
    nl = np.floor(thickness / f32(deltaZ) + f32(0.5)).astype(np.int64)  # nint
    ext = np.zeros((nX, nY, nLayers), f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        sigma = np.where(opticalDepth > np.finfo(f32).tiny, opticalDepth / (nl.astype(f32) * f32(deltaZ)), f32(0)).astype(f32)
    k = np.arange(nLayers)[None, None, :]
    ext = np.where(k < nl[:, :, None], sigma[:, :, None], f32(0)).astype(f32)
    ssa = np.where(ext > 0, f32(SSA), f32(0)).astype(f32)
    pfi = np.where(ext > 0, 1, 0).astype(np.int32)
    addOpticalComponent(d, "cloud", ext, ssa, pfi, _hg_table(0.85, nLegendreCoefficients, "Henyey-Greenstein with g = 0.85"))
    return d


def c1_table():
    """The Deirmendjian C1 phase function, tabulated (i3rcRadarCloud.f95:69-75, 96-100)."""
    dat = np.load(_DATA)
    ang = (dat["c1_angle_deg"].astype(f32) * np.arccos(f32(-1.0)) / f32(180.0)).astype(f32)
    return new_PhaseFunctionTable(ang, dat["c1_value"].astype(f32)[:, None], [1.0], tableDescription="Dermeindjian C1")


def radar_cloud(SSA=1.0, phase="HG", nLegendreCoefficients=299):
    dat = np.load(_DATA)
    nColumns, nLayers, deltaX, deltaZ = 640, 54, f32(50.0), f32(45.0)
    tau = dat["radar_tau"]                       # rows top -> bottom: extinction(:, 1, j), j = nLayers..1
    ext = (tau[::-1, :].T[:, None, :] / deltaZ).astype(f32)  # (x, 1, z)
    d = new_Domain(deltaX * np.arange(0, nColumns + 1, dtype=f32), np.array([0.0, deltaX * nColumns], f32),
                   deltaZ * np.arange(0, nLayers + 1, dtype=f32))
    table = _hg_table(0.85, nLegendreCoefficients) if phase == "HG" else c1_table()
    addOpticalComponent(d, "cloud: " + phase, ext, np.full_like(ext, SSA), np.ones(ext.shape, np.int32), table)
    return d


def synthetic_les(nx=512, ny=512, nz=256, dxy=50.0, dz=20.0, cloud_base=40, cloud_top=120, n_entries=27,
                  mean_tau=12.0, gas_tau=0.3, seed=12345, nLegendreCoefficients=64):
    """C5: LES-like stratocumulus, nC = 2 (cloud + horizontally uniform absorbing gas)."""
    rng = np.random.default_rng(seed)
    cloud_base, cloud_top = int(cloud_base * nz / 256), max(int(cloud_top * nz / 256), int(cloud_base * nz / 256) + 1)
    nzc = cloud_top - cloud_base
    # red-noise (k^-5/3-like) horizontal structure, lognormal amplitude, adiabatic-like vertical profile
    kx = np.fft.fftfreq(nx)[:, None]
    ky = np.fft.rfftfreq(ny)[None, :]
    k = np.sqrt(kx * kx + ky * ky)
    k[0, 0] = 1.0
    amp = k ** (-11.0 / 6.0)
    amp[0, 0] = 0.0
    ext = np.zeros((nx, ny, nzc), f32)
    base = np.fft.irfft2(amp * (rng.standard_normal(amp.shape) + 1j * rng.standard_normal(amp.shape)), s=(nx, ny))
    base = (base - base.mean()) / base.std()
    for kz in range(nzc):
        pert = np.fft.irfft2(amp * (rng.standard_normal(amp.shape) + 1j * rng.standard_normal(amp.shape)), s=(nx, ny))
        pert = (pert - pert.mean()) / pert.std()
        fld = np.exp(0.8 * (0.8 * base + 0.6 * pert))
        top = cloud_top - cloud_base - 12.0 * nz / 256 * (0.5 - 0.5 * np.tanh(base))  # variable cloud top
        prof = (kz + 0.5) / nzc
        ext[:, :, kz] = np.where(kz < top, fld * prof, 0.0).astype(f32)
    ext[ext < 0.15 * ext.mean()] = 0.0                      # holes
    colTau = ext.sum(axis=2) * dz
    ext *= f32(mean_tau / colTau.mean())
    # effective radius grows with height -> phase-function index 1..n_entries
    pfi = np.zeros(ext.shape, np.int32)
    lev = (1 + (np.arange(nzc) * n_entries) // nzc).astype(np.int32)
    pfi[:] = lev[None, None, :]
    pfi[ext <= 0] = 0
    ssa = np.where(ext > 0, f32(0.999), f32(0)).astype(f32)
    gs = np.linspace(0.80, 0.87, n_entries)
    pfs = [new_PhaseFunction((f32(g) ** np.arange(1, nLegendreCoefficients + 1)).astype(f32)) for g in gs]
    table = new_PhaseFunctionTable(pfs, np.linspace(4.0, 17.0, n_entries).astype(f32), tableDescription="synthetic HG by r_eff")
    d = new_Domain(f32(dxy) * np.arange(nx + 1, dtype=f32), f32(dxy) * np.arange(ny + 1, dtype=f32),
                   f32(dz) * np.arange(nz + 1, dtype=f32))
    addOpticalComponent(d, "cloud", ext, ssa, pfi, table, zLevelBase=cloud_base + 1)
    # gas: exponential profile, ssa = 0, horizontally uniform, whole column
    zmid = (np.arange(nz) + 0.5) * dz
    gas = np.exp(-zmid / 2000.0)
    gas = (gas * gas_tau / (gas.sum() * dz)).astype(f32)
    addOpticalComponent(d, "gas", gas, np.zeros(nz, f32), np.ones(nz, np.int32),
                        _hg_table(0.0, 2, "isotropic placeholder"))
    return d
