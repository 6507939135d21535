"""The data formats either side of the hot path: netCDF-3 domain and phase-function-table files
laid out like Code/opticalProperties.f95:554-844 and Code/scatteringPhaseFunctions.f95:899-1252, and the result
files of Example-Drivers/monteCarloDriver.f95:436-854; plus the whole driver on the CPU checker backend."""
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

from i3rc_monte_carlo_model_b200 import fields, fileIO
from i3rc_monte_carlo_model_b200.driver import monteCarloDriver, read_namelists
from i3rc_monte_carlo_model_b200.ErrorMessages import ErrorMessage, getCurrentMessage, stateIsFailure, stateIsSuccess


def _same_domain(a, b):
    for k in ("xPosition", "yPosition", "zPosition"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    assert a.xyRegularlySpaced == b.xyRegularlySpaced and a.zRegularlySpaced == b.zRegularlySpaced
    assert len(a.components) == len(b.components)
    for ca, cb in zip(a.components, b.components):
        assert ca.name == cb.name and ca.zLevelBase == cb.zLevelBase and ca.horizontallyUniform == cb.horizontallyUniform
        assert np.array_equal(ca.extinction, cb.extinction)
        assert np.array_equal(ca.singleScatteringAlbedo, cb.singleScatteringAlbedo)
        assert np.array_equal(ca.phaseFunctionIndex, cb.phaseFunctionIndex)
        ta, tb = ca.table, cb.table
        assert np.array_equal(ta.key, tb.key) and ta.oneAngleSet == tb.oneAngleSet
        for pa, pb in zip(ta.phaseFunctions, tb.phaseFunctions):
            if pa.legendreCoefficients is not None:
                assert np.array_equal(pa.legendreCoefficients, pb.legendreCoefficients)
            else:
                assert np.array_equal(pa.scatteringAngle, pb.scatteringAngle) and np.array_equal(pa.value, pb.value)
            assert pa.extinction == pytest.approx(pb.extinction) and pa.singleScatteringAlbedo == pytest.approx(pb.singleScatteringAlbedo)


@pytest.mark.parametrize("make", [lambda: fields.step_cloud(0.99), lambda: fields.radar_cloud(0.99, "C1"),
                                  lambda: fields.synthetic_les(nx=8, ny=6, nz=10, n_entries=3, seed=2)],
                         ids=["stepCloud-legendre", "radar-angle-value", "les-two-components"])
def test_domain_file_round_trip(tmp_path, make):
    d = make()
    path = str(tmp_path / "domain.dom")
    st = ErrorMessage()
    fileIO.write_Domain(d, path, st)
    assert stateIsSuccess(st), getCurrentMessage(st)
    assert open(path, "rb").read(4) == b"CDF\x01"  # classic netCDF-3, what netCDF-Fortran's nf90_create(nf90_Clobber) writes
    st = ErrorMessage()
    back = fileIO.read_Domain(path, st)
    assert not stateIsFailure(st), getCurrentMessage(st)
    _same_domain(d, back)


def test_domain_file_layout_is_the_reference_layout(tmp_path):
    """Names, types and dimension order as written by write_Domain / add_PhaseFunctionTable through netCDF-Fortran
    (which reverses the dimension order: a(x,y,z) is stored [z][y][x])."""
    d = fields.synthetic_les(nx=8, ny=6, nz=10, n_entries=3, seed=2)
    path = str(tmp_path / "les.dom")
    fileIO.write_Domain(d, path)
    f = netcdf_file(path, "r", mmap=False)
    assert {"x-Edges": 9, "y-Edges": 7, "z-Edges": 11, "x-Grid": 8, "y-Grid": 6, "z-Grid": 10}.items() <= dict(f.dimensions).items()
    assert int(f.numberOfComponents) == len(d.components) == 2
    assert f.xyRegularlySpaced.dtype.itemsize == 1 and int(f.xyRegularlySpaced) == 1
    v = f.variables["Component1_Extinction"]
    c = d.components[0]
    fills = c.zLevelBase == 1 and c.extinction.shape[2] == 10  # else the component brings its own z dimension (:609-616)
    zdim = "z-Grid" if fills else "Component1_z-Grid"
    assert f.dimensions[zdim] == c.extinction.shape[2]
    assert v.dimensions == (zdim, "y-Grid", "x-Grid") and v.data.dtype.kind == "f" and v.data.dtype.itemsize == 4
    assert f.variables["Component1_PhaseFunctionIndex"].data.dtype.itemsize == 2  # nf90_short
    assert np.array_equal(np.array(v[:]), np.transpose(c.extinction, (2, 1, 0)))
    # the second component is horizontally uniform: 1-D variables
    assert len(f.variables["Component2_Extinction"].dimensions) == 1
    assert f.Component1_phaseFunctionStorageType.decode() == "LegendreCoefficients"
    start, length = np.array(f.variables["Component1_start"][:]), np.array(f.variables["Component1_length"][:])
    assert start[0] == 1 and np.array_equal(start[1:], 1 + np.cumsum(length)[:-1])
    assert f.dimensions["Component1_coefficents"] == int(length.sum())
    assert f.Component1_Name.decode() == c.name and int(f.Component1_zLevelBase) == c.zLevelBase
    f.close()


def test_phase_function_table_files(tmp_path):
    t = fields.radar_cloud(0.99, "C1").components[0].table
    path = str(tmp_path / "c1.pft")
    fileIO.write_PhaseFunctionTable(t, path)
    f = netcdf_file(path, "r", mmap=False)
    assert f.phaseFunctionStorageType.decode() == "Angle-Value"
    assert f.variables["phaseFunctionValues"].dimensions == ("phaseFunctionNumber", "scatteringAngle")
    f.close()
    back = fileIO.read_PhaseFunctionTable(path)
    assert back.oneAngleSet and np.array_equal(back.phaseFunctions[0].value, t.phaseFunctions[0].value)
    st = ErrorMessage()
    fileIO.read_PhaseFunctionTable(str(tmp_path / "missing.pft"), status=st)
    assert stateIsFailure(st) and "Can't open file" in getCurrentMessage(st)
    st = ErrorMessage()
    fileIO.read_Domain(path, st)  # a table file is not a domain file
    assert stateIsFailure(st) and "doesn't look an optical properties file" in getCurrentMessage(st)


def test_fortran_edit_descriptors():
    assert fileIO._F(0.5, 7, 3) == "  0.500" and fileIO._F(-0.25, 9, 4) == "  -0.2500" and fileIO._F(12.3456, 5, 2) == "12.35"
    assert fileIO._F(0.85, 5, 2) == " 0.85" and fileIO._F(1234567.0, 7, 3) == "*******"
    assert fileIO._E13_6(1.0) == " 0.100000E+01" and fileIO._E13_6(1365.5) == " 0.136550E+04" and fileIO._E13_6(0.0123) == " 0.123000E-01"


NML = """
! a comment
&radiativeTransfer
  solarFlux = 1., solarMu = .5, solarAzimuth = 0., surfaceAlbedo = 0.2,
  intensityMus  = 1., .5, .5, 0., 0.
  intensityPhis = 0., 0., 180.
/
&monteCarlo
  numPhotonsPerBatch = 2000, numBatches = 4, iseed = 10, nPhaseIntervals = 10001
/
&algorithms
  useRayTracing = .true., useRussianRoulette = .true.,
  useRussianRouletteForIntensity = .true., zetaMin = 0.3
/
&output
  reportAbsorptionProfile = .true., reportVolumeAbsorption = .false.
/
&fileNames
  domainFileName = "{dom}",
  outputFluxFile = "{out}/flux.txt", outputRadFile = "{out}/rad.txt", outputAbsProfFile = "{out}/prof.txt",
  outputNetcdfFile = "{out}/results.nc"
/
"""


def check_driver_outputs(tmp_path, stats):
    flux = open(tmp_path / "flux.txt").read().splitlines()
    assert flux[0] == "!   I3RC Monte Carlo 3D Solar Radiative Transfer: Flux"
    assert flux[2] == "!  Num_Photons=      8000" and flux[3] == "!  PhotonTracing=T    Russian_Roulette=T"
    assert flux[5] == "!  Solar_Flux= 0.100000E+01   Solar_Mu= 0.5000000   Solar_Phi=  0.000"
    assert flux[6] == "!  Lambertian_Surface_Albedo= 0.2000" and flux[7] == "!  Output_Type= Pixel Flux"
    assert flux[8] == "!  Upwelling_Level=250.000   Downwelling_level=  0.000"
    avg = flux[11]
    assert avg.startswith("!  Average:   ")
    nums = [float(t) for t in avg[14:].split()]
    assert nums[0] == pytest.approx(float(stats["meanFluxUp"][0]), abs=6e-5) and len(nums) == 6
    rows = [ln for ln in flux[12:]]
    assert len(rows) == 32 and rows[0][:14] == "  7.812250.000"  # F7.3 x-centre, F7.3 y-centre (no separator, as written)
    # closure from the file itself: up + absorbed + (1 - albedo) * down = 1 within noise
    assert nums[0] + nums[4] + 0.8 * nums[2] == pytest.approx(1.0, abs=0.02)
    rad = open(tmp_path / "rad.txt").read().splitlines()
    assert rad[5].startswith("!  Intensity_uses_Russian_Roulette=T   Intensity_Russian_Roulette_zeta_min= 0.30")
    assert rad[10] == "!  RADIANCE AT Z=250.000   NXO=  32   NYO=   1   NDIR=   3"
    assert rad[12] == "!   1.00000   0.00  <- (mu,phi)" and len(rad) == 12 + 3 * 33
    prof = open(tmp_path / "prof.txt").read().splitlines()
    assert len(prof) == 10 + 32 and prof[7] == "!  Output_Type= Absorption Profile"
    f = netcdf_file(str(tmp_path / "results.nc"), "r", mmap=False)
    assert f.description.decode() == "Output from I3RC Community Monte Carlo Model"
    assert int(f.Total_number_of_photons) == 8000 and int(f.Number_of_batches) == 4 and f.Algorithm.decode() == "Ray_tracing"
    assert f.variables["fluxUp"].dimensions == ("y", "x") and f.variables["intensity"].dimensions == ("direction", "y", "x")
    assert np.allclose(np.array(f.variables["fluxUp"][:]).T, stats["fluxUp"][0], atol=1e-6)
    assert np.allclose(np.array(f.variables["intensity_StdErr"][:]).T, stats["radiance"][1], atol=1e-6)
    assert np.allclose(f.variables["x"][:], 15.625 * (np.arange(32) + 0.5)) and "absorptionProfile" in f.variables
    assert "absorbedVolume" not in f.variables and np.allclose(f.variables["intensityMus"][:], [1.0, 0.5, 0.5])
    f.close()


def test_monteCarloDriver_on_the_checker_backend(tmp_path, oracle):
    dom = str(tmp_path / "step.dom")
    fileIO.write_Domain(fields.step_cloud(0.99), dom)
    nml = tmp_path / "run.nml"
    nml.write_text(NML.format(dom=dom, out=str(tmp_path)))
    cfg = read_namelists(str(nml))
    assert cfg["radiativetransfer"]["intensityMus"] == [1.0, 0.5, 0.5] and cfg["filenames"]["domainFileName"] == dom
    stats = monteCarloDriver(str(nml), backend=oracle, verbose=False)
    check_driver_outputs(tmp_path, stats)


@pytest.mark.gpu
def test_monteCarloDriver_cuda(tmp_path, cuda, oracle):
    dom = str(tmp_path / "step.dom")
    fileIO.write_Domain(fields.step_cloud(0.99), dom)
    nml = tmp_path / "run.nml"
    nml.write_text(NML.format(dom=dom, out=str(tmp_path)))
    stats = monteCarloDriver(str(nml), backend=cuda, verbose=False)
    check_driver_outputs(tmp_path, stats)
    ref = monteCarloDriver(str(nml), backend=oracle, verbose=False)
    for k in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanRadiance"):
        z = (np.asarray(stats[k][0]) - np.asarray(ref[k][0])) / np.sqrt(np.asarray(stats[k][1]) ** 2 + np.asarray(ref[k][1]) ** 2 + 1e-12)
        assert np.all(np.abs(z) < 4.0), (k, z)  # 4 batches only: a loose check, the parity suite does the real one
