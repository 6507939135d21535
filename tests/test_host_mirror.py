"""Host-side mirror of the reference's module interface: argument meaning and error behaviour
(specifyParameters MCRT:872-947, new_PhotonStream checks monteCarloIllumination.f95:78-83,
validateOpticalComponent opticalProperties.f95:929-987, reportResults size checks MCRT:746-791),
exercised through the C-ABI conventions on the CPU checker backend; the `-m gpu` suite repeats the status
checks against the CUDA library."""
import numpy as np
import pytest

from i3rc_monte_carlo_model_b200 import fields
from i3rc_monte_carlo_model_b200.driver import BatchStatistics, partition_batches, read_namelists
from i3rc_monte_carlo_model_b200.ErrorMessages import (ErrorMessage, getCurrentMessage, stateIsFailure, stateIsSuccess,
                                                      stateIsWarning)
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, isReady_Integrator,
                                                                    new_Integrator, reportResults, specifyParameters)
from i3rc_monte_carlo_model_b200.opticalProperties import addOpticalComponent, getInfo_Domain, new_Domain
from i3rc_monte_carlo_model_b200.RandomNumbers import new_RandomNumberSequence
from i3rc_monte_carlo_model_b200.scatteringPhaseFunctions import new_PhaseFunction, new_PhaseFunctionTable
from i3rc_monte_carlo_model_b200.surfaceProperties import new_SurfaceDescription

SPECIFY_CASES = [
    (dict(surfaceAlbedo=1.5), "failure", "surface albedo out of range"),
    (dict(surfaceAlbedo=0.1, surfaceBDRF="uniform"), "failure", "only one surface specification"),
    (dict(minInverseTableSize=100), "warning", "minInverseTableSize less than default"),
    (dict(minForwardTableSize=100), "warning", "minForwardTableSize less than default"),
    (dict(intensityMus=[0.5]), "failure", "Both or neither of intensityMus and intensityPhis"),
    (dict(intensityMus=[1.5], intensityPhis=[0.0]), "failure", "intensityMus must be between -1 and 1"),
    (dict(intensityMus=[0.0], intensityPhis=[0.0]), "failure", "can't be 0"),
    (dict(intensityMus=[0.5], intensityPhis=[400.0]), "failure", "intensityPhis must be between 0 and 360"),
    (dict(computeIntensity=True), "failure", "Can't compute intensity without specifying directions"),
    (dict(zetaMin=-1.0), "warning", "zetaMin must be >= 0"),
    (dict(zetaMin=2.0), "warning", "kind of large"),
    (dict(hybridPhaseFunWidth=45.0), "warning", "hybridPhaseFunWidth out of range"),
    (dict(numOrdersOrigPhaseFunIntenCalcs=-2), "warning", "less than 0"),
    (dict(maxIntensityContribution=0.0), "warning", "maxIntensityContribution <= 0"),
    (dict(surfaceAlbedo=0.3, useRayTracing=False, useRussianRoulette=False), "success", ""),
    (dict(intensityMus=[1.0, -0.5], intensityPhis=[0.0, 90.0], computeIntensity=False), "warning", "Will compute intensity"),
]


def check_specify(backend, kw, state, text):
    status = ErrorMessage()
    I = new_Integrator(fields.plane_parallel(), status=status, backend=backend)
    assert stateIsSuccess(status) and isReady_Integrator(I)
    kw = dict(kw)
    if kw.get("surfaceBDRF") == "uniform":
        kw["surfaceBDRF"] = new_SurfaceDescription([0.2])
    specifyParameters(I, status=status, **kw)
    got = "failure" if stateIsFailure(status) else "warning" if stateIsWarning(status) else "success"
    assert got == state, (kw, getCurrentMessage(status))
    assert text in getCurrentMessage(status)


@pytest.mark.parametrize("kw,state,text", SPECIFY_CASES)
def test_specifyParameters_status(oracle, kw, state, text):
    check_specify(oracle, kw, state, text)


SOURCE_CASES = [
    (dict(solarMu=0.5, solarAzimuth=0.0, numberOfPhotons=0), "non-negative number of photons"),
    (dict(solarMu=1.5, solarAzimuth=0.0, numberOfPhotons=10), "solarMu out of bounds"),
    (dict(solarMu=0.0, solarAzimuth=0.0, numberOfPhotons=10), "solarMu out of bounds"),
    (dict(solarMu=0.5, solarAzimuth=400.0, numberOfPhotons=10), "solarAzimuth out of bounds"),
]


@pytest.mark.parametrize("kw,text", SOURCE_CASES)
def test_new_PhotonStream_status(kw, text):
    status = ErrorMessage()
    new_PhotonStream(status=status, **kw)
    assert stateIsFailure(status) and text in getCurrentMessage(status)


def test_all_six_sources_resolve():
    kinds = [
        (dict(solarMu=0.5, solarAzimuth=10.0), 1), (dict(solarMu=0.5), 2), (dict(), 3),
        (dict(solarMu=0.5, solarAzimuth=0.0, solarX=0.5, solarY=0.25), 4),
        (dict(detectorX=0.5, detectorY=0.5, detectorZ=0.5, detectorPointsUp=True), 5),
        (dict(detectorX=0.5, detectorY=0.5, detectorZ=0.5, detectorMu=-0.5, detectorPhi=10.0), 6),
    ]
    for kw, kind in kinds:
        assert new_PhotonStream(numberOfPhotons=5, **kw).as_c().kind == kind
    ph = new_PhotonStream(0.5, 0.0, numberOfPhotons=3)
    ph.xPosition = ph.yPosition = ph.zPosition = np.full(3, 0.5, np.float32)
    ph.initialMu, ph.initialPhi = np.full(3, -0.5, np.float32), np.zeros(3, np.float32)
    assert ph.as_c().kind == 7 and ph.as_c().numberOfPhotons == 3


def test_domain_validation():
    status = ErrorMessage()
    new_Domain([0.0, 1.0, 1.0], [0.0, 1.0], [0.0, 1.0], status=status)
    assert stateIsFailure(status) and "increasing, unique" in getCurrentMessage(status)
    status = ErrorMessage()
    d = new_Domain([0.0, 1.0, 2.0], [0.0, 1.0], [0.0, 1.0, 2.0, 4.0], status=status)
    assert stateIsSuccess(status) and d.xyRegularlySpaced and not d.zRegularlySpaced
    table = new_PhaseFunctionTable([new_PhaseFunction(np.array([0.8, 0.6], np.float32))], [1.0])
    ok = np.ones((2, 1, 3), np.float32)
    for ext, ssa, pfi, base, text in [
        (-ok, ok, ok.astype(np.int32), 1, "extinction must be >= 0"),
        (ok, 2 * ok, ok.astype(np.int32), 1, "singleScatteringAlbedo must be between 0 and 1"),
        (ok, ok, 3 * ok.astype(np.int32), 1, "phase function index is out of bounds"),
        (ok, ok, ok.astype(np.int32), 3, "vertical extent"),
        (np.ones((3, 1, 3), np.float32), np.ones((3, 1, 3), np.float32), np.ones((3, 1, 3), np.int32), 1, "horizontal extent"),
    ]:
        status = ErrorMessage()
        addOpticalComponent(d, "c", ext, ssa, pfi, table, zLevelBase=base, status=status)
        assert stateIsFailure(status)
        assert any(text in m for _, m in status.messages), (text, status.messages)
    status = ErrorMessage()
    addOpticalComponent(d, "gas", np.ones(3, np.float32), np.zeros(3, np.float32), np.ones(3, np.int32), table, status=status)
    assert stateIsSuccess(status) and d.components[0].horizontallyUniform
    assert getInfo_Domain(d)["numberOfComponents"] == 1


def test_phase_function_validation():
    status = ErrorMessage()
    new_PhaseFunction(np.array([1.5, 0.5], np.float32), status=status)
    assert stateIsFailure(status) and "Asymmetery parameter" in getCurrentMessage(status)
    status = ErrorMessage()
    new_PhaseFunctionTable(np.array([0.0, 1.0, 2.0], np.float32), np.ones((3, 1), np.float32), [1.0], status=status)
    assert stateIsFailure(status) and "Last scattering angle must be max value" in getCurrentMessage(status)


def test_compute_and_report_status(oracle):
    status = ErrorMessage()
    I = new_Integrator(fields.plane_parallel(), status=status, backend=oracle)
    specifyParameters(I, surfaceAlbedo=0.0, status=status)
    computeRadiativeTransfer(I, new_RandomNumberSequence([10, 1]), new_PhotonStream(0.5, 0.0, numberOfPhotons=100), status=status)
    assert stateIsSuccess(status) and "finished with photons" in getCurrentMessage(status)
    r = reportResults(I, "meanIntensity", status=status)
    assert stateIsFailure(status) and "intensity information not available" in getCurrentMessage(status) and r == {}
    status = ErrorMessage()
    r = reportResults(I, "fluxUp", out={"fluxUp": np.zeros((2, 2), np.float32)}, status=status)
    assert stateIsFailure(status) and "wrong size" in getCurrentMessage(status)


def test_partition_and_statistics():
    assert partition_batches(1)[0] == 2  # numBatches = max(numBatches, 2)
    nb, mine = partition_batches(10, 4, 1)
    assert nb == 12 and list(mine) == [4, 5, 6]
    allb = sorted(b for p in range(8) for b in partition_batches(80, 8, p)[1])
    assert allb == list(range(1, 81))
    st = BatchStatistics()
    xs = [1.0, 2.0, 4.0]
    for x in xs:
        st.add({"q": np.array([x, 2 * x])})
    mean, err = st.finish(1.0, 3)["q"]
    assert np.allclose(mean, [np.mean(xs), 2 * np.mean(xs)])
    assert np.allclose(err, [np.std(xs) / np.sqrt(2), 2 * np.std(xs) / np.sqrt(2)])  # sqrt((m2 - m^2)/(nB-1))


def test_namelist_reader(tmp_path):
    p = tmp_path / "mc.nml"
    p.write_text("""
! comment
&radiativeTransfer
  solarFlux = 2., solarMu = 0.5,
  solarAzimuth = 30.  ! trailing
  surfaceAlbedo = 0.1,
  intensityMus = 1., .5, -0.5
  intensityPhis = 0., 0., 180.
/
&monteCarlo
  numPhotonsPerBatch = 1000, numBatches = 8, iseed = 7,
/
&algorithms
  useRayTracing = .false., useRussianRouletteForIntensity = T, zetaMin = 0.25
/
&fileNames
  domainFileName = "a b.dom", outputRadFile = 'r.out'
/
""")
    nml = read_namelists(str(p))
    assert nml["radiativetransfer"]["solarFlux"] == 2.0 and nml["radiativetransfer"]["intensityMus"] == [1.0, 0.5, -0.5]
    assert nml["radiativetransfer"]["intensityPhis"] == [0.0, 0.0, 180.0]
    assert nml["montecarlo"] == dict(numPhotonsPerBatch=1000, numBatches=8, iseed=7, nPhaseIntervals=10001)
    assert nml["algorithms"]["useRayTracing"] is False and nml["algorithms"]["zetaMin"] == 0.25
    assert nml["algorithms"]["useRussianRouletteForIntensity"] is True and nml["algorithms"]["useRussianRoulette"] is True
    assert nml["filenames"]["domainFileName"] == "a b.dom" and nml["filenames"]["outputRadFile"] == "r.out"
    assert nml["output"]["reportVolumeAbsorption"] is False
