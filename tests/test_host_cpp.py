"""The compiled host side (host/): the reference's drivers written in C++ against the C ABI, with their own namelist
reader and netCDF-3 reader/writer.  CPU tests check the readers/writers against the Python mirror file by file; the GPU
tests run the two programs and compare their output with the Python driver's (same device batch loop -> same numbers)."""
import os
import subprocess

import numpy as np
import pytest
from scipy.io import netcdf_file

from i3rc_monte_carlo_model_b200 import fields, fileIO
from tests.test_file_formats import NML, _same_domain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "host", "_build")


@pytest.fixture(scope="module")
def hostbin():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    return lambda name: os.path.join(BUILD, name)


@pytest.mark.parametrize("make", [lambda: fields.step_cloud(0.99), lambda: fields.radar_cloud(0.99, "C1"),
                                  lambda: fields.synthetic_les(nx=8, ny=6, nz=10, n_entries=3, seed=2)],
                         ids=["stepCloud-legendre", "radar-angle-value", "les-two-components"])
def test_cpp_reads_and_rewrites_domain_files(tmp_path, hostbin, make):
    d = make()
    a, b = str(tmp_path / "a.dom"), str(tmp_path / "b.dom")
    fileIO.write_Domain(d, a)
    out = subprocess.run([hostbin("hostio_check"), "domain", a, b], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == f"nx {d.xPosition.size - 1} ny {d.yPosition.size - 1} nz {d.zPosition.size - 1} ncomp {len(d.components)}"
    for line, c in zip(out[1:], d.components):
        assert f"component '{c.name}' zbase {c.zLevelBase} nz {c.extinction.shape[2]} uniform {int(c.horizontallyUniform)}" in line
        sums = line.split("sums")[1].split()
        assert float(sums[0]) == pytest.approx(float(c.extinction.astype(np.float64).sum()), rel=1e-6)
        assert int(sums[2]) == int(c.phaseFunctionIndex.sum())
    assert open(b, "rb").read(4) == b"CDF\x01"
    back = fileIO.read_Domain(b)  # what the C++ writer produced, read by the Python mirror
    _same_domain(d, back)
    fa, fb = netcdf_file(a, "r", mmap=False), netcdf_file(b, "r", mmap=False)
    assert dict(fa.dimensions) == dict(fb.dimensions) and set(fa.variables) == set(fb.variables)
    for k, v in fa.variables.items():
        assert v.dimensions == fb.variables[k].dimensions and v.data.dtype == fb.variables[k].data.dtype, k
    assert fa.xyRegularlySpaced == fb.xyRegularlySpaced and fb.xyRegularlySpaced.dtype.itemsize == 1
    fa.close(), fb.close()


def test_cpp_namelist_reader_and_edit_descriptors(tmp_path, hostbin):
    nml = tmp_path / "run.nml"
    nml.write_text(NML.format(dom="/some/where/step.dom", out="/out"))
    out = subprocess.run([hostbin("hostio_check"), "nml", str(nml)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "radiativeTransfer.solarMu = .5 |"
    assert out[1] == "radiativeTransfer.intensityMus = 1. | .5 | .5 | 0. | 0. |"
    assert out[2] == "radiativeTransfer.intensityPhis = 0. | 0. | 180. |"
    assert out[3] == "monteCarlo.numPhotonsPerBatch = 2000 |" and out[4] == "algorithms.useRayTracing = .true. |"
    assert out[6] == "fileNames.domainFileName = /some/where/step.dom |" and out[7] == "fileNames.outputNetcdfFile = /out/results.nc |"
    assert out[9] == "logical 1 0 real 0.300000"
    fmt = subprocess.run([hostbin("hostio_check"), "fmt"], capture_output=True, text=True, check=True).stdout.splitlines()
    want1 = "".join(f"[{fileIO._F(*a)}]" for a in ((0.5, 7, 3), (-0.25, 9, 4), (12.3456, 5, 2), (0.85, 5, 2), (1234567.0, 7, 3)))
    want2 = "".join(f"[{fileIO._E13_6(v)}]" for v in (1.0, 1365.5, 0.0123))
    assert fmt == [want1, want2]


def test_cpp_drivers_refuse_to_run_without_a_gpu(tmp_path, hostbin, cuda_absent):
    dom = str(tmp_path / "step.dom")
    fileIO.write_Domain(fields.step_cloud(0.99), dom)
    nml = tmp_path / "run.nml"
    nml.write_text(NML.format(dom=dom, out=str(tmp_path)))
    r = subprocess.run([hostbin("monteCarloDriver"), str(nml)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


PP_NML = """
&radiativeTransfer
  solarMu = 0.5, solarAzimuth = 0., surfaceAlbedo = 0.0, {dirs}
/
&monteCarlo
  numPhotonsPerBatch = 20000, numBatches = 4, iseed = 10, nPhaseintervals = 10000
/
&algorithms
  useRayTracing = T, useRussianRoulette = T, useRussianRouletteForIntensity = .false., zetaMin = 0.,
/
&filenames
  domainFileName = "{dom}",
/
&problemOptics
  SSA = 0.95, opticalDepth = 2., g = 0.85, nLegendreCoefficients = 64, useMoments = T,
/
&problemDomain
  nX = 2, nY = 1, domainSize = 500., nLayers = 3, physicalThickness = 250., useSurfaceProperties = F,
/
"""


@pytest.mark.gpu
def test_cpp_monteCarloDriver_matches_the_python_driver(tmp_path, hostbin, cuda):
    from i3rc_monte_carlo_model_b200.driver import monteCarloDriver
    dom = str(tmp_path / "step.dom")
    fileIO.write_Domain(fields.step_cloud(0.99), dom)
    outs = {}
    for who in ("py", "cpp"):
        out = tmp_path / who
        out.mkdir()
        nml = tmp_path / f"{who}.nml"
        nml.write_text(NML.format(dom=dom, out=str(out)))
        if who == "py":
            monteCarloDriver(str(nml), backend=cuda, verbose=False)
        else:
            r = subprocess.run([hostbin("monteCarloDriver"), str(nml)], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            assert "Wrote ASCII results" in r.stdout and "Wrote netcdf results" in r.stdout
        outs[who] = out
    for name in ("flux.txt", "rad.txt", "prof.txt"):  # same device loop, same formats: the text files are identical
        assert open(outs["py"] / name).read() == open(outs["cpp"] / name).read(), name
    fa, fb = netcdf_file(str(outs["py"] / "results.nc"), "r", mmap=False), netcdf_file(str(outs["cpp"] / "results.nc"), "r", mmap=False)
    assert set(fa.variables) == set(fb.variables) and dict(fa.dimensions) == dict(fb.dimensions)
    for k in fa.variables:
        assert fa.variables[k].dimensions == fb.variables[k].dimensions
        # (two runs of the device loop: float32 atomic tallies agree to summation order)
        assert np.allclose(np.array(fa.variables[k][:]), np.array(fb.variables[k][:]), rtol=2e-4, atol=1e-7), k
    for k in ("Total_number_of_photons", "Number_of_batches", "Random_number_seed", "Algorithm", "Solar_mu", "Intensity_uses_Russian_roulette"):
        assert getattr(fa, k) == getattr(fb, k), k
    fa.close(), fb.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dirs", ["", "intensityMus = 1., .5, intensityPhis = 0., 180."], ids=["fluxes", "radiances"])
def test_cpp_planeParallel_matches_the_python_mirror(tmp_path, hostbin, cuda, dirs):
    from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
    from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, new_Integrator, reportResults,
                                                                        specifyParameters)
    from i3rc_monte_carlo_model_b200.RandomNumbers import new_RandomNumberSequence
    dom = str(tmp_path / "pp.dom")
    nml = tmp_path / "pp.nml"
    nml.write_text(PP_NML.format(dom=dom, dirs=dirs))
    r = subprocess.run([hostbin("planeParallel"), str(nml)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0].strip() == f"Wrote domain to file {dom}"
    d = fileIO.read_Domain(dom)  # the domain the C++ program built (createDomain) and wrote
    assert d.xPosition.tolist() == [0.0, 250.0, 500.0] and d.zPosition.size == 4 and len(d.components) == 1
    assert np.allclose(d.components[0].extinction, 2.0 / 250.0) and np.allclose(d.components[0].singleScatteringAlbedo, 0.95)
    I = new_Integrator(d, backend=cuda)
    specifyParameters(I, surfaceAlbedo=0.0)
    if dirs:
        specifyParameters(I, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0])
    specifyParameters(I, useRayTracing=True, useRussianRoulette=True, useHybridPhaseFunsForIntenCalcs=False, hybridPhaseFunWidth=7.0,
                      useRussianRouletteForIntensity=False, zetaMin=0.0)
    per_batch = []
    for b in range(1, 5):
        computeRadiativeTransfer(I, new_RandomNumberSequence([b, 10]), new_PhotonStream(0.5, 0.0, numberOfPhotons=20000))
        res = reportResults(I, "intensity") if dirs else reportResults(I, "fluxUp", "fluxDown", "fluxAbsorbed")
        per_batch.append(res)
    if dirs:
        assert lines[1].split() == "tau omega g theta0 mu phi radiance error".split()
        for i, row in enumerate(lines[2:4]):
            vals = row.split()
            want = np.mean([pb["intensity"][:, :, i] for pb in per_batch])
            assert float(vals[6]) == pytest.approx(want, abs=1.5e-6) and vals[0] == "2.00" and vals[1] == "0.950"
    else:
        vals = lines[2].split()
        for col, key in ((4, "fluxUp"), (5, "fluxDown"), (8, "fluxAbsorbed")):
            assert float(vals[col]) == pytest.approx(np.mean([pb[key] for pb in per_batch]), abs=1.5e-5), key
        assert abs(float(vals[4]) + float(vals[5]) + float(vals[8]) - 1.0) < 0.01


@pytest.mark.gpu
def test_cpp_monteCarloDriver_two_components_with_volume_absorption(tmp_path, hostbin, cuda):
    """A cloud component that does not fill the column (zLevelBase > 1) plus a horizontally uniform gas component, volume
    absorption reported: the C++ driver and the Python driver write the same files."""
    from i3rc_monte_carlo_model_b200.driver import monteCarloDriver
    dom = str(tmp_path / "les.dom")
    fileIO.write_Domain(fields.synthetic_les(nx=12, ny=8, nz=24, n_entries=3, seed=4, nLegendreCoefficients=16), dom)
    nml_text = NML.replace("reportVolumeAbsorption = .false.", "reportVolumeAbsorption = .true.").replace(
        'outputAbsProfFile = "{out}/prof.txt",', 'outputAbsProfFile = "{out}/prof.txt", outputAbsVolumeFile = "{out}/vol.txt",')
    outs = {}
    for who in ("py", "cpp"):
        out = tmp_path / who
        out.mkdir()
        nml = tmp_path / f"{who}.nml"
        nml.write_text(nml_text.format(dom=dom, out=str(out)))
        if who == "py":
            monteCarloDriver(str(nml), backend=cuda, verbose=False)
        else:
            r = subprocess.run([hostbin("monteCarloDriver"), str(nml)], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
        outs[who] = out
    for name in ("flux.txt", "rad.txt", "prof.txt", "vol.txt"):
        a, b = open(outs["py"] / name).read().splitlines(), open(outs["cpp"] / name).read().splitlines()
        assert len(a) == len(b) and a[:8] == b[:8], name
        # numbers: float32 atomic tallies of two runs agree to summation order, i.e. to the last printed digit or so
        for la, lb in zip(a[12:], b[12:]):
            if la.startswith("!"):
                assert la == lb
                continue
            import re  # the F9.4 values (coordinates are F7.3 and written without separators)
            va, vb = (np.array(re.findall(r"-?\d+\.\d{4}(?!\d)", t), float) for t in (la, lb))
            assert va.size == vb.size and va.size >= 2 and np.allclose(va, vb, atol=2.1e-4), (name, la, lb)
    vol = open(outs["cpp"] / "vol.txt").read().splitlines()
    assert len(vol) == 10 + 12 * 8 * 24 and vol[7] == "!  Output_Type= Volume Absorption "
    fb = netcdf_file(str(outs["cpp"] / "results.nc"), "r", mmap=False)
    assert fb.variables["absorbedVolume"].dimensions == ("z", "y", "x") and fb.variables["absorbedVolume"].shape == (24, 8, 12)
    fa = netcdf_file(str(outs["py"] / "results.nc"), "r", mmap=False)
    assert np.allclose(np.array(fa.variables["absorbedVolume"][:]), np.array(fb.variables["absorbedVolume"][:]), rtol=1e-3, atol=1e-7)
    fa.close(), fb.close()
