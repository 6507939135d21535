"""The reference's own shipped namelist files (Example-Drivers/monteCarloDriver.nml, planeParallel.nml), copied verbatim
into tests/golden/reference_namelists/, drive the Python driver and the C++ drivers unchanged."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from i3rc_monte_carlo_model_b200 import fields, fileIO
from i3rc_monte_carlo_model_b200.driver import read_namelists

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NML_DIR = os.path.join(ROOT, "tests", "golden", "reference_namelists")
BUILD = os.path.join(ROOT, "host", "_build")


@pytest.fixture(scope="module")
def hostbin():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    return lambda name: os.path.join(BUILD, name)


def test_python_reader_parses_the_shipped_monteCarloDriver_namelist():
    n = read_namelists(os.path.join(NML_DIR, "monteCarloDriver.nml"))
    rt, mc, al, fn, out = n["radiativetransfer"], n["montecarlo"], n["algorithms"], n["filenames"], n["output"]
    assert (rt["solarFlux"], rt["solarMu"], rt["solarAzimuth"], rt["surfaceAlbedo"]) == (1.0, 0.5, 0.0, 0.0)
    assert rt["intensityMus"] == [1.0, 0.5, 0.5] and rt["intensityPhis"] == [0.0, 0.0, 180.0]
    assert (mc["numPhotonsPerBatch"], mc["numBatches"], mc["iseed"], mc["nPhaseIntervals"]) == (10000, 4, 10, 10001)
    assert al["useRayTracing"] is True and al["useRussianRoulette"] is True and al["useRussianRouletteForIntensity"] is True
    assert al["zetaMin"] == 0.3 and al["useHybridPhaseFunsForIntenCalcs"] is False and al["limitIntensityContributions"] is False
    assert fn["domainFileName"] == "../Tools/Examples/mixture.dom" and fn["outputRadFile"] == "exampleRads.out"
    assert fn["outputFluxFile"] == "exampleFluxes.out" and fn["outputAbsProfFile"] == "exampleAbsorption.out"
    assert fn["outputNetcdfFile"] == "exampleOutput.nc" and not fn["outputAbsVolumeFile"]  # (commented out in the file)
    assert out["reportAbsorptionProfile"] is False and out["reportVolumeAbsorption"] is False


def test_cpp_reader_parses_the_shipped_namelists(hostbin):
    r = subprocess.run([hostbin("hostio_check"), "nml", os.path.join(NML_DIR, "monteCarloDriver.nml")], capture_output=True,
                       text=True, check=True).stdout.splitlines()
    got = dict(line.split(" =", 1) for line in r if " =" in line)
    assert got["radiativeTransfer.solarMu"].split("|")[0].strip() == "0.5"
    assert [t.strip() for t in got["radiativeTransfer.intensityMus"].split("|") if t.strip()] == ["1.", ".5", ".5"]
    assert [t.strip() for t in got["radiativeTransfer.intensityPhis"].split("|") if t.strip()] == ["0.", "0.", "180."]
    assert got["monteCarlo.numPhotonsPerBatch"].split("|")[0].strip() == "10000"
    assert got["fileNames.domainFileName"].split("|")[0].strip().strip('"') == "../Tools/Examples/mixture.dom"
    assert r[-1].startswith("logical 1 0 real 0.300000")
    r = subprocess.run([hostbin("hostio_check"), "nml", os.path.join(NML_DIR, "planeParallel.nml")], capture_output=True,
                       text=True, check=True).stdout.splitlines()
    got = dict(line.split(" =", 1) for line in r if " =" in line)
    assert got["radiativeTransfer.intensityMus"].strip() == ""  # commented out in the shipped file
    assert r[-1].startswith("logical 1 1 real 0.000000")  # useRayTracing = T; the [output] default; zetaMin = 0.


def _tree(tmp_path, domain):
    """The directory layout the shipped namelist assumes: run in Example-Drivers/, domain in ../Tools/Examples/."""
    run = tmp_path / "Example-Drivers"
    (tmp_path / "Tools" / "Examples").mkdir(parents=True)
    run.mkdir()
    fileIO.write_Domain(domain, str(tmp_path / "Tools" / "Examples" / "mixture.dom"))
    shutil.copy(os.path.join(NML_DIR, "monteCarloDriver.nml"), run / "monteCarloDriver.nml")
    return run


@pytest.mark.gpu
def test_monteCarloDriver_runs_from_the_shipped_namelist(tmp_path, hostbin, cuda, oracle):
    """C++ driver and Python driver, started in Example-Drivers/ with the unmodified file: same text outputs; the
    numbers agree with the oracle run of the same namelist (4 batches of 10000 photons) within the family-wise 3-sigma
    bound."""
    from i3rc_monte_carlo_model_b200.driver import monteCarloDriver
    from tests.cases import familywise_bound, make_integrator, oracle_summary
    d = fields.synthetic_les(nx=12, ny=10, nz=16, n_entries=3, seed=4, nLegendreCoefficients=16)  # a two-component "mixture"
    outs = {}
    for who in ("py", "cpp"):
        (tmp_path / who).mkdir()
        run = _tree(tmp_path / who, d)
        cwd = os.getcwd()
        os.chdir(run)
        try:
            if who == "py":
                res = monteCarloDriver("monteCarloDriver.nml", backend=cuda, verbose=False)
            else:
                r = subprocess.run([hostbin("monteCarloDriver"), "monteCarloDriver.nml"], capture_output=True, text=True)
                assert r.returncode == 0, r.stderr + r.stdout
        finally:
            os.chdir(cwd)
        outs[who] = run
        for f in ("exampleRads.out", "exampleFluxes.out", "exampleAbsorption.out", "exampleOutput.nc"):
            assert (run / f).exists(), (who, f)
        assert not (run / "exampleVolumeAbsorption.out").exists()
    for f in ("exampleRads.out", "exampleFluxes.out", "exampleAbsorption.out"):
        # same device loop, same formats: the same text up to the last printed digit (float32 atomic tallies are
        # summed in a different order from run to run)
        a, b = open(outs["py"] / f).read().split(), open(outs["cpp"] / f).read().split()
        assert len(a) == len(b), f
        for ta, tb in zip(a, b):
            if ta != tb:
                last_digit = 10.0 ** -len(ta.split(".")[1].split("E")[0].split("e")[0]) if "." in ta else 1.0
                scale = 10.0 ** int(ta.upper().split("E")[1]) if "E" in ta.upper() else 1.0
                assert abs(float(ta) - float(tb)) <= max(2e-4 * abs(float(ta)), 1.01 * last_digit * scale), (f, ta, tb)
    O = make_integrator(oracle, d, surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0],
                        useRayTracing=True, useRussianRoulette=True, useRussianRouletteForIntensity=True, zetaMin=0.3,
                        minInverseTableSize=10001)
    ref = oracle_summary(O, 10000, 32)
    zs = []
    for k, kr in (("meanFluxUp", "meanFluxUp"), ("meanFluxDown", "meanFluxDown"), ("meanFluxAbsorbed", "meanFluxAbsorbed")):
        m, e = res[k]
        zs.append((float(m) - float(ref[kr][0])) / np.hypot(float(e), float(ref[kr][1])))
    m, e = res["meanRadiance"]
    zs += list((np.ravel(m) - np.ravel(ref["meanIntensity"][0])) / np.hypot(np.ravel(e), np.ravel(ref["meanIntensity"][1])))
    assert np.all(np.abs(zs) <= familywise_bound(len(zs), 3)), zs  # (the driver's 4 batches: 3 degrees of freedom)


@pytest.mark.gpu
def test_planeParallel_runs_from_the_shipped_namelist(tmp_path, hostbin, cuda, oracle):
    """`planeParallel planeParallel.nml` with the unmodified file (no domain file name: nothing is written; no radiance
    directions: the flux line) against the oracle on the same problem (tau = 1, g = 0.85, mu0 = 0.5, black surface)."""
    from tests.cases import familywise_bound, make_integrator, oracle_summary
    shutil.copy(os.path.join(NML_DIR, "planeParallel.nml"), tmp_path / "planeParallel.nml")
    r = subprocess.run([hostbin("planeParallel"), "planeParallel.nml"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr + r.stdout
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    vals = lines[-1].split()
    assert vals[0] == "1.00" and vals[1] == "1.000" and vals[2] == "0.850"
    up, down, absorbed = float(vals[4]), float(vals[5]), float(vals[8])
    O = make_integrator(oracle, fields.plane_parallel(), surfaceAlbedo=0.0)
    ref = oracle_summary(O, 10000, 32)
    se4 = lambda k: float(ref[k][1]) * np.sqrt(32 / 4.0)  # the program ran 4 batches of the same size
    zs = [(up - float(ref["meanFluxUp"][0])) / np.hypot(se4("meanFluxUp"), float(ref["meanFluxUp"][1])),
          (down - float(ref["meanFluxDown"][0])) / np.hypot(se4("meanFluxDown"), float(ref["meanFluxDown"][1]))]
    assert np.all(np.abs(zs) <= familywise_bound(2, 31)), (zs, up, down)
    assert absorbed == 0.0 and abs(up + down - 1.0) < 2e-5
