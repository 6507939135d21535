"""The data formats either side of the hot path, bit-compatible with the files
the reference reads and writes through netCDF-Fortran (classic netCDF-3 format):

  * domain files            read_Domain / write_Domain                 Code/opticalProperties.f95:554-844
  * phase-function tables   add_/read_/write_PhaseFunctionTable        Code/scatteringPhaseFunctions.f95:899-1252
  * results                 writeResults_ASCII / writeResults_netcdf   Example-Drivers/monteCarloDriver.f95:436-854

netCDF-Fortran reverses the dimension order: a Fortran array a(x, y, z) is the file variable a[z, y, x] (x fastest),
a table values(angle, entry) is values[entry, angle].  ``scipy.io.netcdf_file`` writes and reads the classic format.
"""
from __future__ import annotations

import os

import numpy as np
from scipy.io import netcdf_file

from .ErrorMessages import setStateToFailure, setStateToSuccess, setStateToWarning, stateIsFailure
from .opticalProperties import addOpticalComponent, domain, finalize_Domain, new_Domain
from .scatteringPhaseFunctions import (isReady_PhaseFunctionTable, new_PhaseFunction, new_PhaseFunctionTable,
                                       phaseFunctionTable)


def _prefix(i):  # makePrefix, opticalProperties.f95:1006-1016
    return f"Component{i}_"


def _att_str(v):
    return v.decode() if isinstance(v, bytes) else str(v)


# ---- phase-function tables ----------------------------------------------------------------------------------------
def add_PhaseFunctionTable(table: phaseFunctionTable, f: netcdf_file, prefix="", status=None):
    """scatteringPhaseFunctions.f95:928-1114.  ``f`` is an open, writable netcdf_file."""
    if not isReady_PhaseFunctionTable(table):
        setStateToFailure(status, "add_PhaseFunctionTable: phase function table hasn't been initialized.")
        return
    legendre = all(p.legendreCoefficients is not None for p in table.phaseFunctions)
    if not (table.oneAngleSet or legendre):
        setStateToFailure(status, "add_PhaseFunctionTable: Can't write general phase function tables to files.")
        return
    if prefix + "phaseFunctionNumber" in f.dimensions:
        setStateToFailure(status, "add_PhaseFunctionTable: trying to add to the wrong kind of file")
        return
    n = len(table.phaseFunctions)
    f.createDimension(prefix + "phaseFunctionNumber", n)
    f.createVariable(prefix + "phaseFunctionKeyT", "f", (prefix + "phaseFunctionNumber",))[:] = np.asarray(table.key, np.float32)
    f.createVariable(prefix + "extinctionT", "f", (prefix + "phaseFunctionNumber",))[:] = np.array(
        [p.extinction for p in table.phaseFunctions], np.float32)
    f.createVariable(prefix + "singleScatteringAlbedoT", "f", (prefix + "phaseFunctionNumber",))[:] = np.array(
        [p.singleScatteringAlbedo for p in table.phaseFunctions], np.float32)
    if table.description.strip():
        setattr(f, prefix + "description", table.description.strip())
    if table.oneAngleSet:
        ang = np.asarray(table.phaseFunctions[0].scatteringAngle, np.float32)
        f.createDimension(prefix + "scatteringAngle", ang.size)
        f.createVariable(prefix + "scatteringAngle", "f", (prefix + "scatteringAngle",))[:] = ang
        f.createVariable(prefix + "phaseFunctionValues", "f", (prefix + "phaseFunctionNumber", prefix + "scatteringAngle"))[:] = \
            np.stack([np.asarray(p.value, np.float32) for p in table.phaseFunctions])
        setattr(f, prefix + "phaseFunctionStorageType", "Angle-Value")
    else:
        length = np.array([p.legendreCoefficients.size for p in table.phaseFunctions], np.int32)
        start = (np.concatenate([[0], np.cumsum(length)[:-1]]) + 1).astype(np.int32)  # 1-based, like the reference
        f.createDimension(prefix + "coefficents", int(length.sum()))  # (sic: the reference's spelling)
        f.createVariable(prefix + "start", "i", (prefix + "phaseFunctionNumber",))[:] = start
        f.createVariable(prefix + "length", "i", (prefix + "phaseFunctionNumber",))[:] = length
        f.createVariable(prefix + "legendreCoefficients", "f", (prefix + "coefficents",))[:] = np.concatenate(
            [np.asarray(p.legendreCoefficients, np.float32) for p in table.phaseFunctions])
        setattr(f, prefix + "phaseFunctionStorageType", "LegendreCoefficients")
    setStateToSuccess(status)


def write_PhaseFunctionTable(table, fileName, status=None):
    """scatteringPhaseFunctions.f95:899-926"""
    f = netcdf_file(fileName, "w")
    try:
        add_PhaseFunctionTable(table, f, "", status)
    finally:
        f.close()


def read_PhaseFunctionTable(fileName=None, f=None, prefix="", status=None) -> phaseFunctionTable:
    """scatteringPhaseFunctions.f95:1116-1252: from a file name, or from an open file with a prefix."""
    own = f is None
    if own:
        try:
            f = netcdf_file(fileName, "r", mmap=False)
        except (OSError, TypeError, ValueError):
            setStateToFailure(status, f"read_PhaseFunctionTable: Can't open file {fileName}")
            return phaseFunctionTable()
    try:
        try:
            kind = _att_str(getattr(f, prefix + "phaseFunctionStorageType"))
            key = np.array(f.variables[prefix + "phaseFunctionKeyT"][:], np.float32)
            ext = np.array(f.variables[prefix + "extinctionT"][:], np.float32)
            ssa = np.array(f.variables[prefix + "singleScatteringAlbedoT"][:], np.float32)
        except (AttributeError, KeyError):
            setStateToFailure(status, "read_PhaseFunctionTable: file doesn't look like a phase function table.")
            return phaseFunctionTable()
        desc = _att_str(getattr(f, prefix + "description", ""))
        if kind.strip() == "Angle-Value":
            ang = np.array(f.variables[prefix + "scatteringAngle"][:], np.float32)
            vals = np.array(f.variables[prefix + "phaseFunctionValues"][:], np.float32)  # [entry, angle]
            return new_PhaseFunctionTable(ang, vals.T.copy(), key, extinction=ext, singleScatteringAlbedo=ssa,
                                          tableDescription=desc, status=status)
        if kind.strip() == "LegendreCoefficients":
            start = np.array(f.variables[prefix + "start"][:], np.int64)
            length = np.array(f.variables[prefix + "length"][:], np.int64)
            coefs = np.array(f.variables[prefix + "legendreCoefficients"][:], np.float32)
            pfs = [new_PhaseFunction(coefs[s - 1:s - 1 + n], extinction=float(e), singleScatteringAlbedo=float(a))
                   for s, n, e, a in zip(start, length, ext, ssa)]
            return new_PhaseFunctionTable(pfs, key, tableDescription=desc, status=status)
        setStateToFailure(status, "read_PhaseFunctionTable: Unknown phase function table format.")
        return phaseFunctionTable()
    finally:
        if own:
            f.close()


# ---- domains ------------------------------------------------------------------------------------------------------
def write_Domain(thisDomain: domain, fileName, status=None):
    """opticalProperties.f95:554-706"""
    if thisDomain.xPosition is None:
        setStateToFailure(status, "write_Domain: domain hasn't been initialized.")
        return
    f = netcdf_file(fileName, "w")
    ok = True
    try:
        nx, ny, nz = (a.size - 1 for a in (thisDomain.xPosition, thisDomain.yPosition, thisDomain.zPosition))
        for name, n in (("x-Edges", nx + 1), ("y-Edges", ny + 1), ("z-Edges", nz + 1), ("x-Grid", nx), ("y-Grid", ny),
                        ("z-Grid", nz)):
            f.createDimension(name, n)
        f.createVariable("x-Edges", "f", ("x-Edges",))[:] = thisDomain.xPosition
        f.createVariable("y-Edges", "f", ("y-Edges",))[:] = thisDomain.yPosition
        f.createVariable("z-Edges", "f", ("z-Edges",))[:] = thisDomain.zPosition
        f.xyRegularlySpaced = np.int8(thisDomain.xyRegularlySpaced)  # asInt: one byte
        f.zRegularlySpaced = np.int8(thisDomain.zRegularlySpaced)
        if thisDomain.components:
            f.numberOfComponents = np.int32(len(thisDomain.components))
        for i, c in enumerate(thisDomain.components, 1):
            p = _prefix(i)
            setattr(f, p + "Name", c.name.strip())
            setattr(f, p + "zLevelBase", np.int32(c.zLevelBase))
            ncz = c.extinction.shape[2]
            if c.zLevelBase == 1 and ncz == nz:
                zdim = "z-Grid"
            else:
                zdim = p + "z-Grid"
                f.createDimension(zdim, ncz)
            dims = (zdim,) if c.horizontallyUniform else (zdim, "y-Grid", "x-Grid")
            for vname, arr, code in (("Extinction", c.extinction, "f"), ("SingleScatteringAlbedo", c.singleScatteringAlbedo, "f"),
                                     ("PhaseFunctionIndex", c.phaseFunctionIndex, "h")):
                a = arr[0, 0, :] if c.horizontallyUniform else np.transpose(arr, (2, 1, 0))
                f.createVariable(p + vname, code, dims)[:] = a.astype(np.int16 if code == "h" else np.float32)
            add_PhaseFunctionTable(c.table, f, p, status)
            if status is not None and stateIsFailure(status):
                ok = False
                break
    finally:
        f.close()
    if not ok:
        os.remove(fileName)  # the reference deletes a half-written file (:693-700)
        return
    setStateToSuccess(status)


def read_Domain(fileName, status=None) -> domain:
    """opticalProperties.f95:708-844"""
    try:
        f = netcdf_file(fileName, "r", mmap=False)
    except (OSError, TypeError, ValueError):
        setStateToFailure(status, f"read_Domain: Can't open file {fileName}")
        return domain()
    try:
        try:
            x, y, z = (np.array(f.variables[k][:], np.float32) for k in ("x-Edges", "y-Edges", "z-Edges"))
            if "z-Grid" not in f.dimensions:
                raise KeyError("z-Grid")
        except KeyError:
            setStateToFailure(status, f"read_Domain: {fileName} doesn't look an optical properties file.")
            return domain()
        d = new_Domain(x, y, z, status)
        if status is not None and stateIsFailure(status):
            return d
        if bool(int(np.ravel(getattr(f, "xyRegularlySpaced", 0))[0])) != d.xyRegularlySpaced:
            setStateToWarning(status, "read_Domain: file and new domain don't agree on regularity of x-y spacing.")
        if bool(int(np.ravel(getattr(f, "zRegularlySpaced", 0))[0])) != d.zRegularlySpaced:
            setStateToWarning(status, "read_Domain: file and new domain don't agree on regularity of z spacing.")
        ncomp = int(np.ravel(getattr(f, "numberOfComponents", 0))[0])
        for i in range(1, ncomp + 1):
            p = _prefix(i)
            try:
                name = _att_str(getattr(f, p + "Name"))
                zbase = int(np.ravel(getattr(f, p + "zLevelBase"))[0])
                v = f.variables[p + "Extinction"]
                uniform = len(v.dimensions) == 1
                fields = []
                for vname, dt in (("Extinction", np.float32), ("SingleScatteringAlbedo", np.float32), ("PhaseFunctionIndex", np.int32)):
                    a = np.array(f.variables[p + vname][:])
                    a = a.reshape(1, 1, -1) if uniform else np.transpose(a, (2, 1, 0))
                    fields.append(np.asfortranarray(a.astype(dt)))
            except (AttributeError, KeyError):
                setStateToFailure(status, f"read_Domain: Error reading scalar fields from file {fileName}")
                return d
            table = read_PhaseFunctionTable(f=f, prefix=p, status=status)
            if status is not None and stateIsFailure(status):
                setStateToFailure(status, "read_Domain: Error reading phase function table.")
                return d
            addOpticalComponent(d, name, fields[0], fields[1], fields[2], table, zLevelBase=zbase, status=status)
            if status is not None and stateIsFailure(status):
                return d
        setStateToSuccess(status)
        return d
    finally:
        f.close()


# ---- results (monteCarloDriver.f95:436-854) -------------------------------------------------------------------------
def _L(b):
    return "T" if b else "F"


def _E13_6(v):
    """Fortran E13.6: 0.ddddddE+ee, right-justified in 13 columns."""
    if v == 0:
        return " 0.000000E+00"
    e = int(np.floor(np.log10(abs(v)))) + 1
    m = v / 10.0**e
    if abs(round(m, 6)) >= 1.0:
        m /= 10.0
        e += 1
    return f"{m:9.6f}E{e:+03d}".rjust(13)


def _F(v, w, d):
    s = f"{float(v):{w}.{d}f}"
    if s.startswith("0.") and len(s) > w:  # Fortran drops the optional leading zero when the field is full
        s = s[1:]
    if s.startswith("-0.") and len(s) > w:
        s = "-" + s[2:]
    return s if len(s) <= w else "*" * w


def _header(fh, title, cfg, out_type, radiance=False):
    fh.write(f"!   I3RC Monte Carlo 3D Solar Radiative Transfer: {title}\n")
    fh.write("!  Property_File=" + f"{cfg['domainFileName'][:60]:<60s}" + "\n")  # A60 of a blank-padded character(256)
    fh.write("!  Num_Photons=" + f"{cfg['numPhotonsPerBatch'] * cfg['numBatches']:10d}" + "\n")
    fh.write(f"!  PhotonTracing={_L(cfg['useRayTracing'])}    Russian_Roulette={_L(cfg['useRussianRoulette'])}\n")
    fh.write(f"!  Hybrid_Phase_Func_for_Radiance={_L(cfg['useHybridPhaseFunsForIntenCalcs'])}"
             f"   Gaussian_Phase_Func_Width_deg={_F(cfg['hybridPhaseFunWidth'], 5, 2)}\n")
    if radiance:
        fh.write(f"!  Intensity_uses_Russian_Roulette={_L(cfg['useRussianRouletteForIntensity'])}"
                 f"   Intensity_Russian_Roulette_zeta_min={_F(cfg['zetaMin'], 5, 2)}\n")
        fh.write(f"!  limited_intensity_contributions={_L(cfg['limitIntensityContributions'])}"
                 f"   max_intensity_contribution={_F(cfg['maxIntensityContribution'], 5, 2)}\n")
    fh.write(f"!  Solar_Flux={_E13_6(cfg['solarFlux'])}   Solar_Mu={_F(cfg['solarMu'], 10, 7)}"
             f"   Solar_Phi={_F(cfg['solarAzimuth'], 7, 3)}\n")
    fh.write(f"!  Lambertian_Surface_Albedo={_F(cfg['surfaceAlbedo'], 7, 4)}\n")
    fh.write(f"!  Output_Type= {out_type}\n")


def _pair(ms):
    return "".join(" " + _F(v, 9, 4) for v in ms)


def writeResults_ASCII(cfg, xPosition, yPosition, zPosition, stats):
    """monteCarloDriver.f95:436-605.  ``cfg``: the namelist values (flat dict, the reference's names); ``stats``:
    name -> (mean, stderr) arrays indexed [x, y(, z | direction)] (the *Stats arrays of the driver)."""
    x, y, z = (np.asarray(a, np.float64) for a in (xPosition, yPosition, zPosition))
    nx, ny, nz = x.size - 1, y.size - 1, z.size - 1
    xm, ym, zm = (x[:-1] + x[1:]) / 2, (y[:-1] + y[1:]) / 2, (z[:-1] + z[1:]) / 2
    if cfg.get("outputFluxFile"):
        with open(cfg["outputFluxFile"], "w") as fh:
            _header(fh, "Flux", cfg, "Pixel Flux")
            fh.write(f"!  Upwelling_Level={_F(z[nz], 7, 3)}   Downwelling_level={_F(z[0], 7, 3)}\n")
            fh.write("!   X      Y           Flux_Up             Flux_Down            Flux_Absorbed \n")
            fh.write("!                  Mean     StdErr       Mean     StdErr       Mean     StdErr\n")
            fh.write(f"{'!  Average:   ':>14s}" + "".join(" " + _pair((float(stats[k][0]), float(stats[k][1])))
                                                      for k in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed")) + "\n")
            up, dn, ab = stats["fluxUp"], stats["fluxDown"], stats["fluxAbsorbed"]
            for j in range(ny):
                for i in range(nx):
                    fh.write(_F(xm[i], 7, 3) + _F(ym[j], 7, 3) + "".join(" " + _pair((s[0][i, j], s[1][i, j])) for s in (up, dn, ab)) + "\n")
    if cfg.get("outputAbsProfFile"):
        with open(cfg["outputAbsProfFile"], "w") as fh:
            _header(fh, "Absorption Profile", cfg, "Absorption Profile")
            fh.write("!   Z    Absorbed_Flux (flux/km) \n")
            fh.write("!          Mean     StdErr \n")
            m, s = stats["absorbedProfile"]
            for k in range(nz):
                fh.write(_F(zm[k], 7, 3) + " " + _pair((m[k], s[k])) + "\n")
    if cfg.get("outputAbsVolumeFile"):
        with open(cfg["outputAbsVolumeFile"], "w") as fh:
            _header(fh, "3D Absorption Field", cfg, "Volume Absorption ")
            fh.write("!    X       Y        Z       Absorbed_Flux (flux/km)\n")
            fh.write("!                               Mean     StdErr \n")
            m, s = stats["absorbedVolume"]
            for i in range(nx):
                for j in range(ny):
                    for k in range(nz):
                        fh.write(_F(xm[i], 7, 3) + " " + _F(ym[j], 7, 3) + " " + _F(zm[k], 7, 3) + " " + _pair((m[i, j, k], s[i, j, k])) + "\n")
    if cfg.get("outputRadFile"):
        mus, phis = cfg["intensityMus"], cfg["intensityPhis"]
        nd = sum(1 for v in mus if abs(v) > 0)
        with open(cfg["outputRadFile"], "w") as fh:
            _header(fh, "Radiance", cfg, "Pixel Radiance", radiance=True)
            fh.write(f"!  RADIANCE AT Z={_F(z[nz], 7, 3)}   NXO={nx:4d}   NYO={ny:4d}   NDIR={nd:4d}\n")
            fh.write("!   X      Y         Radiance (Mean, StdErr)\n")
            m, s = stats["radiance"]
            for k in range(nd):
                fh.write(f"!  {_F(mus[k], 8, 5)} {_F(phis[k], 6, 2)}  <- (mu,phi)\n")
                for j in range(ny):
                    for i in range(nx):
                        fh.write(_F(xm[i], 7, 3) + _F(ym[j], 7, 3) + _pair((m[i, j, k], s[i, j, k])) + "\n")


def writeResults_netcdf(cfg, xPosition, yPosition, zPosition, stats, cpuTimeTotal=0.0, cpuTimeSetup=0.0, numProcs=1):
    """monteCarloDriver.f95:609-854: the same attributes, dimensions and variables."""
    x, y, z = (np.asarray(a, np.float32) for a in (xPosition, yPosition, zPosition))
    nx, ny, nz = x.size - 1, y.size - 1, z.size - 1
    f = netcdf_file(cfg["outputNetcdfFile"], "w")
    try:
        f.description = "Output from I3RC Community Monte Carlo Model"
        f.Domain_filename = cfg["domainFileName"].strip()
        f.Surface_albedo = np.float32(cfg["surfaceAlbedo"])
        f.Total_number_of_photons = np.int32(cfg["numPhotonsPerBatch"] * cfg["numBatches"])
        f.Number_of_batches = np.int32(cfg["numBatches"])
        f.Solar_flux = np.float32(cfg["solarFlux"])
        f.Solar_mu = np.float32(cfg["solarMu"])
        f.Solar_phi = np.float32(cfg["solarAzimuth"])
        f.Random_number_seed = np.int32(cfg["iseed"])
        f.Phase_function_table_sizes = np.int32(cfg["nPhaseIntervals"])
        f.Algorithm = "Ray_tracing" if cfg["useRayTracing"] else "Max_cross_section"
        hyb = bool(cfg["useHybridPhaseFunsForIntenCalcs"])
        f.Intensity_uses_hyrbid_phase_functions = np.int32(hyb)  # (sic)
        f.Hybrid_phase_function_width = np.float32(cfg["hybridPhaseFunWidth"] if hyb else 0.0)
        rr = bool(cfg["useRussianRouletteForIntensity"])
        f.Intensity_uses_Russian_roulette = np.int32(rr)
        f.Intensity_Russian_roulette_zeta_min = np.float32(cfg["zetaMin"] if rr else 0.0)
        lim = bool(cfg["limitIntensityContributions"])
        f.limited_intensity_contributions = np.int32(lim)
        f.max_intensity_contribution = np.float32(cfg["maxIntensityContribution"] if lim else 0.0)
        f.Cpu_time_total = np.float32(cpuTimeTotal)
        f.Cpu_time_setup = np.float32(cpuTimeSetup)
        f.Number_of_processors_used = np.int32(numProcs)
        prof, vol = bool(cfg.get("reportAbsorptionProfile")), bool(cfg.get("reportVolumeAbsorption"))
        f.createDimension("x", nx)
        f.createDimension("y", ny)
        if prof or vol:
            f.createDimension("z", nz)
        f.createVariable("x", "f", ("x",))[:] = (x[:-1] + x[1:]) / 2
        f.createVariable("y", "f", ("y",))[:] = (y[:-1] + y[1:]) / 2
        if prof or vol:
            f.createVariable("z", "f", ("z",))[:] = (z[:-1] + z[1:]) / 2

        def put(name, a, dims):  # a indexed Fortran-like [x, y, ..] -> file order reversed
            f.createVariable(name, "f", dims)[:] = np.transpose(np.asarray(a, np.float32))

        for k in ("fluxUp", "fluxDown", "fluxAbsorbed"):
            put(k, stats[k][0], ("y", "x"))
        for k in ("fluxUp", "fluxDown", "fluxAbsorbed"):
            put(k + "_StdErr", stats[k][1], ("y", "x"))
        if prof:
            put("absorptionProfile", stats["absorbedProfile"][0], ("z",))
            put("absorptionProfile_StdErr", stats["absorbedProfile"][1], ("z",))
        if vol:
            put("absorbedVolume", stats["absorbedVolume"][0], ("z", "y", "x"))
            put("absorbedVolume_StdErr", stats["absorbedVolume"][1], ("z", "y", "x"))
        if "radiance" in stats:
            nd = np.asarray(stats["radiance"][0]).shape[2]
            f.createDimension("direction", nd)
            f.createVariable("intensityMus", "f", ("direction",))[:] = np.asarray(cfg["intensityMus"][:nd], np.float32)
            f.createVariable("intensityPhis", "f", ("direction",))[:] = np.asarray(cfg["intensityPhis"][:nd], np.float32)
            put("intensity", stats["radiance"][0], ("direction", "y", "x"))
            put("intensity_StdErr", stats["radiance"][1], ("direction", "y", "x"))
    finally:
        f.close()


__all__ = ["add_PhaseFunctionTable", "write_PhaseFunctionTable", "read_PhaseFunctionTable", "write_Domain", "read_Domain",
           "writeResults_ASCII", "writeResults_netcdf", "finalize_Domain"]
