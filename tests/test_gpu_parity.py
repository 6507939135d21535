"""Parity tests proper: the CUDA path, called through the C ABI (include/i3rc_b200.h), against the CPU oracle on
the same inputs and namelists-equivalents.

Criteria (BASELINE.json north_star): because the RNG differs (per-photon Philox streams instead of the shared
MT19937 stream), domain-mean and per-column fluxes, absorption and intensities agree within 3 sigma of the
combined Monte Carlo batch standard error; the deterministic sub-paths (phase-table inversion, optical-path
integration along a fixed ray) agree to 1e-5 relative.
"""
import ctypes as C
import os

import numpy as np
import pytest

from i3rc_monte_carlo_model_b200 import _abi, fields
from i3rc_monte_carlo_model_b200.driver import (allreduce_device_stats, device_stats_report, partition_batches,
                                               run_batches_device, run_batches_host)
from i3rc_monte_carlo_model_b200.ErrorMessages import ErrorMessage, getCurrentMessage, stateIsFailure
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, copy_Integrator, getCounters,
                                                                    getTable, new_Integrator, new_Integrator_dense,
                                                                    reportResults, specifyParameters, traceRays)
from i3rc_monte_carlo_model_b200.RandomNumbers import new_RandomNumberSequence
from i3rc_monte_carlo_model_b200.surfaceProperties import new_SurfaceDescription
from tests.cases import (assert_counter_parity, assert_statistical_parity, make_integrator, mean_se, oracle_summary,
                         run_batches)
from tests.golden.make_golden import CASES as GOLDEN_CASES
from tests.test_host_mirror import SPECIFY_CASES, check_specify

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- deterministic sub-paths ------------------------------------------------------------------------------------
def _rel(a, b, floor):
    return np.max(np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), floor))


@pytest.mark.parametrize("name,make,params", [
    ("HG-64", lambda: fields.plane_parallel(), {}),
    ("HG-299", lambda: fields.landsat_cloud(1.0), dict(minInverseTableSize=10001, minForwardTableSize=10001)),
    ("HG-tabulated", lambda: fields.plane_parallel(useMoments=False, nAngles=5000), {}),
    ("C1-tabulated", lambda: fields.radar_cloud(1.0, "C1"), {}),
    ("multi-entry", lambda: fields.synthetic_les(nx=8, ny=8, nz=16, n_entries=5, seed=3), {}),
])
def test_phase_tables_match_oracle(cuda, oracle, name, make, params):
    """tabulateInverse/ForwardPhaseFunctions on the device vs the oracle: 1e-5 relative."""
    d = make()
    kw = dict(surfaceAlbedo=0.0, intensityMus=[1.0], intensityPhis=[0.0], **params)
    G, O = make_integrator(cuda, d, **kw), make_integrator(oracle, d, **kw)
    assert cuda.tabulate(G.handle) == 0, G._msg()
    oracle.tabulate(O.handle)
    for c in range(G.nc):
        inv_g, inv_o = getTable(G, 0, c), getTable(O, 0, c).astype(np.float64)
        assert inv_g.shape == inv_o.shape
        # angles: 1e-5 relative to pi-scale (entries near 0 are acos() of numbers next to 1: absolute floor 1e-4 rad,
        # the float32 resolution of acos there)
        assert np.max(np.abs(inv_g - inv_o) / np.maximum(inv_o, 10.0)) < 1e-5, name
        assert _rel(np.cos(inv_g), np.cos(inv_o), 1.0) < 1e-5, name
        fwd_g, fwd_o = getTable(G, 2, c), getTable(O, 2, c).astype(np.float64)
        assert _rel(fwd_g, fwd_o, np.abs(fwd_o).max() * 1e-3 + 1e-3) < 1e-5 * 5, name
        assert np.median(np.abs(fwd_g - fwd_o) / np.maximum(np.abs(fwd_o), 1e-3)) < 1e-5, name


def test_hybrid_tables_match_oracle(cuda, oracle):
    d = fields.radar_cloud(1.0, "C1")
    kw = dict(surfaceAlbedo=0.0, intensityMus=[1.0], intensityPhis=[0.0], useHybridPhaseFunsForIntenCalcs=True,
              hybridPhaseFunWidth=7.0)
    G, O = make_integrator(cuda, d, **kw), make_integrator(oracle, d, **kw)
    assert cuda.tabulate(G.handle) == 0
    oracle.tabulate(O.handle)
    hg, ho = getTable(G, 1), getTable(O, 1).astype(np.float64)
    og = getTable(G, 2)
    kg, ko = np.nonzero(hg[0] != og[0])[0].max(), np.nonzero(ho[0] != getTable(O, 2)[0])[0].max()
    assert abs(int(kg) - int(ko)) <= 1  # same transition index (float32 sums may shift it by one)
    assert _rel(hg, ho, 1e-2) < 2e-4


def test_scattering_angle_and_phase_lookup_probes(cuda, oracle):
    """computeScatteringAngle / lookUpPhaseFuncValsFromTable with IDENTICAL (injected) tables: float32 round-off only."""
    d = fields.plane_parallel()
    kw = dict(surfaceAlbedo=0.0, intensityMus=[1.0], intensityPhis=[0.0])
    G, O = make_integrator(cuda, d, **kw), make_integrator(oracle, d, **kw)
    oracle.tabulate(O.handle)
    inv, fwd = getTable(O, 0), getTable(O, 1)
    assert cuda.set_inverse_table(G.handle, 0, inv.shape[1], 1, _abi.fptr(inv)) == 0
    assert cuda.set_forward_table(G.handle, 0, fwd.shape[1], 1, _abi.fptr(fwd), None) == 0
    rng = np.random.default_rng(0)
    xi = np.concatenate([rng.random(5000), [0.0, 1.0, 0.5]]).astype(np.float32)
    tg, to = np.zeros_like(xi), np.zeros_like(xi)
    assert cuda.sample_scattering_angles(G.handle, 0, 0, xi.size, _abi.fptr(xi), _abi.fptr(tg)) == 0
    oracle.sample_scattering_angles(O.handle, 0, 0, xi.size, _abi.fptr(xi), _abi.fptr(to))
    assert np.max(np.abs(tg - to)) < 1e-6
    ang = np.concatenate([rng.random(5000) * np.pi, [0.0, np.pi]]).astype(np.float32)
    pg, po = np.zeros_like(ang), np.zeros_like(ang)
    assert cuda.lookup_phase_function(G.handle, 0, 0, 1, ang.size, _abi.fptr(ang), _abi.fptr(pg)) == 0
    oracle.lookup_phase_function(O.handle, 0, 0, 1, ang.size, _abi.fptr(ang), _abi.fptr(po))
    assert np.max(np.abs(pg - po) / np.abs(po)) < 1e-5


@pytest.fixture
def force_layer_split(monkeypatch):
    """Store only the horizontally varying layers of totalExt in 3-D whatever the size of the field (the library does
    it on its own for fields that do not fit L2)."""
    monkeypatch.setenv("I3RC_SPLIT_LAYERS", "2")


def _irregular_domain_with_uniform_layers():
    d = _irregular_domain()
    c = d.components[0]
    ext, ssa, pfi = c.extinction.copy(), c.singleScatteringAlbedo.copy(), c.phaseFunctionIndex.copy()
    for k, v in ((0, 0.004), (1, 0.0), (6, 0.0), (12, 0.002), (13, 0.0)):  # clear or hazy layers, horizontally uniform
        ext[:, :, k], ssa[:, :, k], pfi[:, :, k] = v, (0.95 if v > 0 else 0.0), (1 if v > 0 else 0)
    from i3rc_monte_carlo_model_b200.opticalProperties import replaceOpticalComponent
    replaceOpticalComponent(d, 1, c.name, ext, ssa, pfi, c.table)
    return d


@pytest.mark.parametrize("field", ["stepCloud", "landsat", "irregular", "uniformLayers", "irregularUniformLayers"])
def test_optical_path_along_fixed_rays(cuda, oracle, field, force_layer_split):
    """accumulateExtinctionAlongPath on the device vs exact float64 integration (1e-5) and vs the oracle."""
    from tests.hostsim.binding import dense_from_domain
    from tests.test_oracle_pins import _f64_optical_path
    if field == "stepCloud":
        d = fields.step_cloud(1.0)
    elif field == "landsat":
        d = fields.landsat_cloud(1.0, nLegendreCoefficients=8)
    elif field == "uniformLayers":  # cloud in 20 of 64 layers + gas everywhere: only the cloudy layers are stored in 3-D
        d = fields.synthetic_les(nx=24, ny=16, nz=64, n_entries=3, seed=7, nLegendreCoefficients=8)
    elif field == "irregularUniformLayers":
        d = _irregular_domain_with_uniform_layers()
    else:
        d = _irregular_domain()
    tot = dense_from_domain(d)[0].astype(np.float64)
    G, O = make_integrator(cuda, d, surfaceAlbedo=0.0), make_integrator(oracle, d, surfaceAlbedo=0.0)
    rng = np.random.default_rng(11)
    n = 300
    lo = np.array([d.xPosition[0], d.yPosition[0], d.zPosition[0]], np.float64)
    hi = np.array([d.xPosition[-1], d.yPosition[-1], d.zPosition[-1]], np.float64)
    pos = (lo + (0.02 + 0.96 * rng.random((n, 3))) * (hi - lo)).astype(np.float32)
    mu = rng.uniform(0.15, 1.0, n) * rng.choice([-1, 1], n)
    phi = rng.uniform(0, 2 * np.pi, n)
    u = np.column_stack([np.sqrt(1 - mu**2) * np.cos(phi), np.sqrt(1 - mu**2) * np.sin(phi), mu]).astype(np.float32)
    tau, _, idx = traceRays(G, pos, u)
    exact = np.array([_f64_optical_path(d, tot, pos[i], u[i]) for i in range(n)])
    big = exact > 1e-2
    assert np.max(np.abs(tau[big] - exact[big]) / exact[big]) < 1e-5
    tau_o, _, idx_o = traceRays(O, pos, u)
    tol_oracle = 5e-5 if field == "stepCloud" else 2e-3  # the reference's absolute float32 positions (see CPU test)
    assert np.max(np.abs(tau[big] - tau_o[big]) / exact[big]) < tol_oracle
    assert np.array_equal(idx[:, 2], idx_o[:, 2])
    # with an optical-path limit: same end cell, same end point
    lim = (exact * rng.uniform(0.05, 0.9, n)).astype(np.float32)
    tg, pg, ig = traceRays(G, pos, u, lim)
    to, po, io = traceRays(O, pos, u, lim)
    assert np.allclose(tg, lim, rtol=1e-6) and np.allclose(to, lim, rtol=1e-6)
    same = np.all(ig == io, axis=1)
    assert same.mean() > 0.97  # end points within an ulp of a cell face may be attributed to either cell
    L = hi - lo
    dpos = np.abs(pg[same] - po[same])
    dpos[:, :2] = np.minimum(dpos[:, :2], L[:2] - dpos[:, :2])
    # end points: the oracle carries absolute float32 positions, good to ~1e-2 m after a long path on the large domains
    assert dpos.max() < (2e-3 if field == "stepCloud" else 0.1)


def _irregular_domain():
    """Irregular x, y and z spacing exercises findIndex-style location and per-cell widths (MCRT:1371-1372, 1386)."""
    from i3rc_monte_carlo_model_b200.opticalProperties import addOpticalComponent, new_Domain
    rng = np.random.default_rng(4)
    x = np.concatenate([[0.0], np.cumsum(rng.uniform(20, 80, 12))]).astype(np.float32)
    y = np.concatenate([[0.0], np.cumsum(rng.uniform(30, 60, 9))]).astype(np.float32)
    z = np.concatenate([[0.0], np.cumsum(rng.uniform(10, 50, 14))]).astype(np.float32)
    d = new_Domain(x, y, z)
    assert not d.xyRegularlySpaced and not d.zRegularlySpaced
    ext = (rng.random((12, 9, 14)) ** 3 * 0.02).astype(np.float32)
    ext[rng.random(ext.shape) < 0.3] = 0
    pfi = np.where(ext > 0, 1, 0).astype(np.int32)
    addOpticalComponent(d, "cloud", ext, np.where(ext > 0, 0.95, 0).astype(np.float32), pfi, fields._hg_table(0.8, 32))
    return d


# ---- statistical parity against the oracle ------------------------------------------------------------------------
STAT_CASES = {
    "C1-planeParallel-flux": (lambda: fields.plane_parallel(), dict(surfaceAlbedo=0.0), dict(solarMu=0.5, solarAzimuth=0.0), 8000),
    "C1-planeParallel-radiance-nml": (lambda: fields.plane_parallel(), dict(
        surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0],
        useRussianRouletteForIntensity=False), dict(solarMu=0.5, solarAzimuth=0.0), 40000),
    "C2-stepCloud-conservative-radiance": (lambda: fields.step_cloud(1.0), dict(
        surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0],
        useRussianRouletteForIntensity=True, zetaMin=0.3), dict(solarMu=0.5, solarAzimuth=0.0), 20000),
    "C2-stepCloud-absorbing-overhead-sun": (lambda: fields.step_cloud(0.99), dict(
        surfaceAlbedo=0.2, intensityMus=[1.0, -0.5], intensityPhis=[0.0, 90.0], useRussianRouletteForIntensity=False),
        dict(solarMu=1.0, solarAzimuth=0.0), 8000),
    "max-cross-section": (lambda: fields.step_cloud(0.99), dict(surfaceAlbedo=0.3, useRayTracing=False),
                          dict(solarMu=0.5, solarAzimuth=0.0), 3000),
    "irregular-grid-radiance": (_irregular_domain, dict(
        surfaceAlbedo=0.4, intensityMus=[0.8], intensityPhis=[45.0], useRussianRouletteForIntensity=True, zetaMin=0.2),
        dict(solarMu=0.7, solarAzimuth=200.0), 20000),
    "two-components-hybrid-limited": (lambda: fields.synthetic_les(nx=16, ny=16, nz=24, n_entries=4, seed=5), dict(
        surfaceAlbedo=0.1, intensityMus=[1.0], intensityPhis=[0.0], useHybridPhaseFunsForIntenCalcs=True,
        hybridPhaseFunWidth=7.0, numOrdersOrigPhaseFunIntenCalcs=1, limitIntensityContributions=True,
        maxIntensityContribution=0.5, useRussianRouletteForIntensity=False), dict(solarMu=0.6, solarAzimuth=20.0), 10000),
    "no-roulette-tabulated": (lambda: fields.plane_parallel(useMoments=False, SSA=0.9, nX=3, nY=2, nLayers=4), dict(
        surfaceAlbedo=0.5, useRussianRoulette=False), dict(solarMu=0.5, solarAzimuth=0.0), 4000),
    # 16 directions: every event batch overfills the warp's task ring, so batches are suspended and resumed
    "sixteen-directions-plain": (lambda: fields.step_cloud(0.99), dict(
        surfaceAlbedo=0.2, intensityMus=[m for m in (1.0, 0.7, 0.4, -0.6) for _ in range(4)],
        intensityPhis=[p for _ in range(4) for p in (0.0, 90.0, 180.0, 270.0)], useRussianRouletteForIntensity=False),
        dict(solarMu=0.5, solarAzimuth=0.0), 6000),
    "sixteen-directions-roulette": (lambda: fields.step_cloud(1.0), dict(
        surfaceAlbedo=0.1, intensityMus=[m for m in (1.0, 0.8, 0.5, 0.3) for _ in range(4)],
        intensityPhis=[p for _ in range(4) for p in (10.0, 100.0, 190.0, 280.0)], useRussianRouletteForIntensity=True,
        zetaMin=0.5), dict(solarMu=0.8, solarAzimuth=40.0), 10000),
    "irregular-grid-uniform-layers": (_irregular_domain_with_uniform_layers, dict(
        surfaceAlbedo=0.4, intensityMus=[0.8, -0.6], intensityPhis=[45.0, 200.0], useRussianRouletteForIntensity=False),
        dict(solarMu=0.7, solarAzimuth=200.0), 20000),
    "source-random-azimuth": (lambda: fields.step_cloud(0.99), dict(surfaceAlbedo=0.0), dict(solarMu=0.6), 3000),
    "source-flux": (lambda: fields.step_cloud(0.99), dict(surfaceAlbedo=0.0), dict(), 3000),
    "source-spotlight": (lambda: fields.step_cloud(0.99), dict(surfaceAlbedo=0.0),
                         dict(solarMu=0.5, solarAzimuth=30.0, solarX=0.7, solarY=0.5), 3000),
    "source-internal-flux": (lambda: fields.step_cloud(0.99), dict(surfaceAlbedo=0.3),
                             dict(detectorX=0.3, detectorY=0.5, detectorZ=0.53, detectorPointsUp=False), 3000),
    "source-internal-intensity": (lambda: fields.step_cloud(0.99), dict(surfaceAlbedo=0.3),
                                  dict(detectorX=0.3, detectorY=0.5, detectorZ=0.6, detectorMu=0.7, detectorPhi=1.0), 3000),
}


@pytest.mark.parametrize("name", ["irregular-grid-uniform-layers", "two-components-hybrid-limited", "sixteen-directions-roulette"])
def test_statistical_parity_with_layer_split(cuda, oracle, name, force_layer_split):
    test_statistical_parity_with_oracle(cuda, oracle, name)


@pytest.mark.parametrize("name", list(STAT_CASES))
def test_statistical_parity_with_oracle(cuda, oracle, name):
    make, params, source, nph = STAT_CASES[name]
    d = make()
    nb = 32
    got = run_batches(make_integrator(cuda, d, **params), nph, nb, source=source)
    ref = oracle_summary(make_integrator(oracle, d, **params), nph, nb, source=source)
    assert_statistical_parity(got, ref, label=name + ": ")
    gc, rc = got["counters"], ref["counters"]
    assert gc["photons"] == nph and gc["bad"] == 0
    assert_counter_parity(got, rc, ("collisions", "surface_hits", "exits_top", "crossings_photon"), label=name + ": ")


def test_surface_brdf_map(cuda, oracle):
    """specifyParameters(surfaceBDRF = ...) with a Lambertian albedo map (surfaceProperties.f95:60-162)."""
    d = fields.plane_parallel(nX=4, nY=2, opticalDepth=0.3)
    xs, ys = np.array([0.0, 250.0, 500.0], np.float32), np.array([0.0, 500.0], np.float32)
    res = {}
    for name, be in (("g", cuda), ("o", oracle)):
        surf = new_SurfaceDescription(np.array([[[0.8], [0.1]]], np.float32), xs, ys)
        I = make_integrator(be, d, surfaceBDRF=surf, intensityMus=[1.0], intensityPhis=[0.0], useRussianRouletteForIntensity=False)
        res[name] = run_batches(I, 4000, 32, source=dict(solarMu=0.8, solarAzimuth=0.0))
    assert_statistical_parity(res["g"], res["o"], label="brdf: ")
    up = res["g"]["fluxUp"].mean(0)
    assert up[:2].mean() > 1.1 * up[2:].mean()  # the bright half reflects more


def test_photon_arrays_source(cuda, oracle):
    """The public array components of type(photonStream) (monteCarloIllumination.f95:34-41) filled by hand."""
    d = fields.step_cloud(0.99)
    rng = np.random.default_rng(1)
    n = 5000
    res = {}
    for name, be in (("g", cuda), ("o", oracle)):
        I = make_integrator(be, d, surfaceAlbedo=0.0)
        vals = []
        for b in range(32):
            r2 = np.random.default_rng(100 + b)
            ph = new_PhotonStream(0.5, 0.0, numberOfPhotons=n)
            ph.xPosition, ph.yPosition = r2.random(n, dtype=np.float32), r2.random(n, dtype=np.float32)
            ph.zPosition = np.full(n, 1.0 - np.finfo(np.float32).eps, np.float32)
            ph.initialMu, ph.initialPhi = np.full(n, -0.5, np.float32), np.zeros(n, np.float32)
            computeRadiativeTransfer(I, new_RandomNumberSequence([10, b + 1]), ph)
            r = reportResults(I, "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxUp")
            vals.append(r)
        res[name] = {k: np.stack([np.asarray(v[k], np.float64) for v in vals]) for k in vals[0]}
    assert_statistical_parity(res["g"], res["o"], label="arrays: ")


# ---- golden fixtures (oracle outputs at larger sizes) ------------------------------------------------------------
def _coarsen(a, f=8):
    nx, ny = a.shape[0], a.shape[1]
    if nx % f or ny % f:
        return a
    return a.reshape(nx // f, f, ny // f, f, *a.shape[2:]).mean(axis=(1, 3))


# GPU photons per batch and batches for each fixture (default: twice the oracle's photons per batch, 16 batches)
GOLDEN_GPU_SIZE = {"landsat_rr_hi": (16_000_000, 64), "les_mid_split": (1_000_000, 16), "radar_c1_rr": (4_000_000, 16),
                   "step_mu1_rr": (8_000_000, 16)}


def _load_golden(name):
    """name -> dict of arrays; a fixture with a second half (name_b: further batches of the same case, independent
    streams) is combined with it: mean of the two means, standard errors in quadrature."""
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    pb = os.path.join(GOLDEN, name + "_b.npz")
    if os.path.exists(pb):
        b = np.load(pb)
        for k in list(g):
            if k.endswith("_mean"):
                g[k] = 0.5 * (g[k].astype(np.float64) + b[k])
            elif k.endswith("_se"):
                g[k] = 0.5 * np.hypot(g[k].astype(np.float64), b[k])
            elif k.startswith("cnt_") or k == "numBatches":
                g[k] = g[k] + b[k]
    g["files"] = list(g)
    return g


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if not n.endswith("_b")])
def test_against_golden_fixture(cuda, name):
    """The CUDA path against committed oracle outputs (tests/golden/make_golden.py).  Domain means: the family-wise
    3-sigma bound of tests/cases.py (FLAT 3 sigma per quantity for landsat_rr_hi, the bench.py workload, where 6.4e7
    oracle photons and 1e9 GPU photons give sigma(meanFluxUp) ~ 1e-4); per-column fields in blocks of 8x8 columns."""
    from tests.cases import familywise_bound
    make, params, source, nph, nb = GOLDEN_CASES[name]
    g = _load_golden(name)
    nbo = int(g["numBatches"])
    I = make_integrator(cuda, make(), **params)
    if name == "les_mid_split":  # large enough for new_Integrator to choose the layer-compacted field by itself
        assert cuda.get_layout(I.handle, 0) > 0, "expected the SPLIT kernel (layer-compacted extinction field)"
    nphg, nbg = GOLDEN_GPU_SIZE.get(name, (nph * 2, 16))
    got = run_batches(I, nphg, nbg, source=source)
    dof = min(nbg, nbo) - 1
    flat = name == "landsat_rr_hi"
    means = ["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed"]
    zs = []
    for k in means:
        m, s = mean_se(got[k])
        zs.append((m - g[k + "_mean"]) / np.sqrt(s**2 + g[k + "_se"] ** 2 + 1e-16))
    bound = 3.0 if flat else familywise_bound(len(zs), dof)
    assert np.all(np.abs(zs) <= bound), (name, means, zs, bound)
    m, s = mean_se(got["absorbedProfile"])
    z = (m - g["absorbedProfile_mean"]) / np.sqrt(s**2 + g["absorbedProfile_se"] ** 2 + 1e-20)
    assert np.abs(z).max() <= familywise_bound(z.size, dof), (name, "absorbedProfile", np.abs(z).max())
    f = int(g["coarsen"]) if "coarsen" in g["files"] else 1  # (fixture already stored as block means)

    def blocks(per_batch, gm, gs):  # per_batch [nb, x, y(, d)], fixture [(d,) y, x] -> z per 8x8 block
        gm, gs = np.moveaxis(gm.T, -1, -1), gs.T
        fb = 8 // f if (per_batch.shape[1] % 8 == 0 and per_batch.shape[2] % 8 == 0) else 1
        m, s = mean_se(np.stack([_coarsen(b, 8 if fb * f == 8 else 1) for b in per_batch]))
        if fb > 1:
            gm_c, gs_c = _coarsen(gm, fb), np.sqrt(_coarsen(gs**2, fb) / fb**2)
        else:
            gm_c, gs_c = gm, gs
        return (m - gm_c) / np.sqrt(s**2 + gs_c**2 + 1e-20)

    if "meanRadiance_mean" in g["files"]:
        m, s = mean_se(got["meanIntensity"])
        z = (m - g["meanRadiance_mean"]) / np.hypot(s, g["meanRadiance_se"])
        assert np.all(np.abs(z) <= (3.0 if flat else familywise_bound(z.size, dof))), (name, "meanRadiance", z)
        z = blocks(got["intensity"], g["radiance_mean"], g["radiance_se"])
        assert np.abs(z).max() <= familywise_bound(z.size, dof) and np.mean(z**2) < 1.6, (name, "radiance", np.abs(z).max(), np.mean(z**2))
    z = blocks(got["fluxUp"], g["fluxUp_mean"], g["fluxUp_se"])
    assert np.abs(z).max() <= familywise_bound(z.size, dof) and np.mean(z**2) < 1.6, (name, "fluxUp", np.abs(z).max(), np.mean(z**2))
    ref_cnt = {k[4:]: float(g[k]) for k in g["files"] if k.startswith("cnt_")}
    assert_counter_parity(got, ref_cnt, ("collisions", "crossings_photon", "exits_top"), label=name + ": ")


# ---- size-independent properties at benchmark size ---------------------------------------------------------------
def test_properties_at_landsat_benchmark_size(cuda):
    """Energy closure, column absorption = integral of the volume absorption, reproducibility per (seed, batch),
    independence of the launch shape, and agreement of the device batch loop with the host loop."""
    d = fields.landsat_cloud(0.99)
    params = dict(surfaceAlbedo=0.3, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0], useRussianRouletteForIntensity=True,
                  zetaMin=0.3)
    I = make_integrator(cuda, d, **params)
    nph = 2_000_000
    r = run_batches(I, nph, 4, want=["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxAbsorbed", "volumeAbsorption",
                                       "fluxUp", "intensity", "meanIntensity"])
    closure = r["meanFluxUp"] + r["meanFluxAbsorbed"] + 0.7 * r["meanFluxDown"]
    assert np.all(np.abs(closure - 1.0) < 2e-3)
    dz = np.diff(d.zPosition)
    col = (r["volumeAbsorption"][0] * dz[None, None, :]).sum(-1)
    assert np.allclose(col, r["fluxAbsorbed"][0], rtol=5e-4, atol=1e-5)
    assert r["counters"]["photons"] == nph and r["counters"]["bad"] == 0
    # same (iseed, batch) -> same photons: results agree to float32 atomic-summation order
    again = run_batches(I, nph, 1, want=["meanFluxUp", "fluxUp", "intensity"])
    assert abs(again["meanFluxUp"][0] - r["meanFluxUp"][0]) < 2e-6
    assert np.allclose(again["fluxUp"][0], r["fluxUp"][0], rtol=2e-4, atol=1e-5)
    # ... independently of the launch shape (per-photon counter-based streams)
    I2 = copy_Integrator(I)
    assert cuda.set_tuning(I2.handle, b"resident_blocks", 4) == 0 and cuda.set_tuning(I2.handle, b"steps_per_event_phase", 8) == 0 and cuda.set_tuning(I2.handle, b"min_running", 24) == 0
    other = run_batches(I2, nph, 1, want=["meanFluxUp", "fluxUp", "meanIntensity"])
    assert abs(other["meanFluxUp"][0] - r["meanFluxUp"][0]) < 2e-6
    assert np.allclose(other["meanIntensity"][0], r["meanIntensity"][0], rtol=1e-4)
    assert other["counters"]["collisions"] == again["counters"]["collisions"]


def test_device_batch_loop_equals_host_loop(cuda):
    """i3rc_run_batches + device moments (monteCarloDriver.f95:300-378) == the same loop through reportResults."""
    d = fields.step_cloud(0.99)
    params = dict(surfaceAlbedo=0.2, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0], useRussianRouletteForIntensity=True)
    src = dict(solarMu=0.5, solarAzimuth=0.0)
    nB, mine = partition_batches(6, 1, 0)
    I = make_integrator(cuda, d, **params)
    host = run_batches_host(I, src, 20000, mine, iseed=10).finish(2.0, nB)
    run_batches_device(I, src, 20000, mine, iseed=10, with_volume=True)
    allreduce_device_stats(I, None)
    dev = device_stats_report(I, 2.0, nB, with_volume=True)
    pairs = [("meanFluxUp", "meanFluxUp"), ("fluxUp", "fluxUp"), ("fluxAbsorbed", "fluxAbsorbed"),
             ("absorbedProfile", "absorbedProfile"), ("intensity", "radiance"), ("meanIntensity", "meanRadiance")]
    for hk, dk in pairs:
        hm, he = host[hk]
        dm, de = dev[dk]
        assert np.allclose(np.squeeze(dm), np.squeeze(hm), rtol=3e-4, atol=1e-6), hk
        assert np.allclose(np.squeeze(de), np.squeeze(he), rtol=2e-2, atol=1e-5), hk


def test_dense_and_component_constructors_agree(cuda):
    """new_Integrator from dense arrays == new_Integrator(domain) with the expansion done on the device."""
    from tests.hostsim.binding import dense_from_domain
    d = fields.synthetic_les(nx=16, ny=16, nz=24, n_entries=4, seed=5)
    tot, cum, ssa, pfi = dense_from_domain(d)
    A = make_integrator(cuda, d, surfaceAlbedo=0.1)
    B = new_Integrator_dense(d.xPosition, d.yPosition, d.zPosition, tot, cum, ssa, pfi, [c.table for c in d.components], backend=cuda)
    specifyParameters(B, surfaceAlbedo=0.1)
    ra = run_batches(A, 20000, 2)
    rb = run_batches(B, 20000, 2)
    assert np.allclose(ra["fluxUp"], rb["fluxUp"], rtol=2e-4, atol=1e-5)
    assert ra["counters"]["collisions"] == rb["counters"]["collisions"]


# ---- edge cases and status behaviour on the CUDA library ------------------------------------------------------
def test_float32_tallies_do_not_saturate_on_a_single_column(cuda):
    """64 M photons into ONE column: float32 tallies alone would stop growing at 2^24 increments (closure 0.43); the
    batch is traced in pieces that are folded into float64 sums."""
    I = make_integrator(cuda, fields.plane_parallel(SSA=0.9), surfaceAlbedo=0.3, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0],
                        useRussianRouletteForIntensity=False)
    small = run_batches(I, 1_000_000, 4, want=["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity", "absorbedProfile"])
    big = run_batches(I, 64_000_000, 1, want=["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity", "absorbedProfile"])
    closure = big["meanFluxUp"][0] + big["meanFluxAbsorbed"][0] + 0.7 * big["meanFluxDown"][0]
    assert abs(closure - 1.0) < 5e-4
    assert big["counters"]["photons"] == 64_000_000
    for k in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity", "absorbedProfile"):
        m, s = mean_se(small[k])
        assert np.all(np.abs(big[k][0] - m) < 5.0 * s + 2e-4), (k, big[k][0], m, s)


@pytest.mark.parametrize("nph", [1, 31, 33, 1000])
def test_ragged_photon_counts(cuda, nph):
    I = make_integrator(cuda, fields.plane_parallel(), surfaceAlbedo=0.0)
    r = run_batches(I, nph, 2)
    assert r["counters"]["photons"] == nph
    assert np.allclose(r["meanFluxUp"] + r["meanFluxDown"], 1.0, atol=1e-6)


@pytest.mark.parametrize("kw,state,text", SPECIFY_CASES)
def test_specifyParameters_status_cuda(cuda, kw, state, text):
    check_specify(cuda, kw, state, text)


def test_compute_status_cuda(cuda):
    status = ErrorMessage()
    I = new_Integrator(fields.plane_parallel(), status=status, backend=cuda)
    ph = new_PhotonStream(0.5, 0.0, numberOfPhotons=10)
    ph.source.solarMu = 2.0
    computeRadiativeTransfer(I, new_RandomNumberSequence([1, 2]), ph, status=status)
    assert stateIsFailure(status) and "solarMu out of bounds" in getCurrentMessage(status)
    status = ErrorMessage()
    r = reportResults(I, "intensity", status=status)
    assert stateIsFailure(status) and "intensity information not available" in getCurrentMessage(status)


# ---- spectral loop: a gas component's extinction profile swapped on the device ------------------
@pytest.mark.gpu
def test_component_profile_swap_equals_a_fresh_integrator(cuda, oracle, force_layer_split):
    from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import setComponentProfile
    from i3rc_monte_carlo_model_b200.opticalProperties import replaceOpticalComponent
    params = dict(surfaceAlbedo=0.1, intensityMus=[1.0, 0.6], intensityPhis=[0.0, 90.0], useRussianRouletteForIntensity=True, zetaMin=0.3)
    d1 = fields.synthetic_les(nx=16, ny=12, nz=32, n_entries=3, seed=3, nLegendreCoefficients=16)
    gas = d1.components[1]
    prof2 = (gas.extinction[0, 0, :] * np.linspace(3.0, 0.5, 32)).astype(np.float32)  # another k-term
    d2 = fields.synthetic_les(nx=16, ny=12, nz=32, n_entries=3, seed=3, nLegendreCoefficients=16)
    replaceOpticalComponent(d2, 2, gas.name, prof2, gas.singleScatteringAlbedo[0, 0, :], gas.phaseFunctionIndex[0, 0, :], gas.table)
    A = make_integrator(cuda, d1, **params)
    setComponentProfile(A, 2, prof2)          # swapped on the device
    B = make_integrator(cuda, d2, **params)   # built from the modified domain
    rng = np.random.default_rng(5)
    n = 200
    hi = np.array([d1.xPosition[-1], d1.yPosition[-1], d1.zPosition[-1]], np.float64)
    pos = ((0.02 + 0.96 * rng.random((n, 3))) * hi).astype(np.float32)
    mu = rng.uniform(0.2, 1.0, n) * rng.choice([-1, 1], n)
    phi = rng.uniform(0, 2 * np.pi, n)
    u = np.column_stack([np.sqrt(1 - mu**2) * np.cos(phi), np.sqrt(1 - mu**2) * np.sin(phi), mu]).astype(np.float32)
    ta, _, _ = traceRays(A, pos, u)
    tb, _, _ = traceRays(B, pos, u)
    assert np.allclose(ta, tb, rtol=2e-6, atol=1e-7)
    ra, rb = run_batches(A, 20000, 8), run_batches(B, 20000, 8)  # same seeds, same photons up to float rounding of the field
    for k in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity"):
        assert np.allclose(ra[k].mean(0), rb[k].mean(0), rtol=5e-3, atol=2e-4), k
    assert abs(ra["counters"]["collisions"] - rb["counters"]["collisions"]) < 0.01 * rb["counters"]["collisions"]
    ref = oracle_summary(make_integrator(oracle, d2, **params), 20000, 16)
    got = run_batches(A, 20000, 16)
    assert_statistical_parity(got, ref, keys=["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity"], label="swapped profile: ")
    st = ErrorMessage()
    setComponentProfile(A, 5, prof2, status=st)
    assert stateIsFailure(st) and "no such component" in getCurrentMessage(st)


@pytest.mark.gpu
def test_spectral_band_loop(cuda):
    from i3rc_monte_carlo_model_b200.driver import run_batches_device, device_stats_report, run_spectral_bands
    d = fields.synthetic_les(nx=16, ny=12, nz=32, n_entries=3, seed=3, nLegendreCoefficients=16)
    gas = d.components[1].extinction[0, 0, :].copy()
    I = make_integrator(cuda, d, surfaceAlbedo=0.1, intensityMus=[1.0], intensityPhis=[0.0])
    src = dict(solarMu=0.6, solarAzimuth=10.0)
    bands = [(0.5, gas * 0.2), (0.3, gas * 1.0), (0.2, gas * 6.0)]
    tot = run_spectral_bands(I, src, 40000, 8, 2, bands, iseed=3)
    # the same three terms one by one, combined by hand
    want = 0.0
    absorbed = []
    for k, (w, prof) in enumerate(bands):
        from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import setComponentProfile
        setComponentProfile(I, 2, prof)
        run_batches_device(I, src, 40000, range(1 + 8 * k, 9 + 8 * k), iseed=3)
        st = device_stats_report(I, 1.0, 8)
        want = want + w * st["meanFluxUp"][0]
        absorbed.append(float(st["meanFluxAbsorbed"][0]))
    assert float(tot["meanFluxUp"][0]) == pytest.approx(float(want), rel=1e-6)
    assert absorbed[0] < absorbed[1] < absorbed[2]  # more gas, more absorption
    assert 0 < float(tot["meanFluxUp"][1]) < 5e-3


@pytest.mark.parametrize("case", ["stepCloud", "lesTwoComponents"])
def test_translation_invariance_of_the_periodic_domain(cuda, case):
    """Shifting a periodic domain shifts the per-column fluxes and radiances (indexing, wrap-around, exit columns)."""
    from tests.cases import assert_translation_invariance
    if case == "stepCloud":
        d, k = fields.step_cloud(0.99), (5, 0)
    else:
        d, k = fields.synthetic_les(nx=16, ny=12, nz=24, n_entries=3, seed=9, nLegendreCoefficients=16), (7, 5)
    assert_translation_invariance(cuda, d, k[0], k[1],
                                  dict(surfaceAlbedo=0.2, intensityMus=[1.0, 0.5, -0.7], intensityPhis=[0.0, 200.0, 90.0],
                                       useRussianRouletteForIntensity=False),
                                  100000, 16, source=dict(solarMu=0.5, solarAzimuth=30.0))


# ---- round 2: device probes of the random stream and of next_direct, C-ABI validation, staged tallies ---------------
def test_philox_on_the_device_matches_known_answers_and_numpy():
    """Philox4x32-10 as the transport kernel runs it (csrc/philox.cuh on the GPU): the Random123 known-answer vector
    that fits the kernel's counter layout (photon id, block, 0), and an independent numpy implementation (itself
    checked against all three known answers, tests/test_philox_numpy.py) for random keys, photons and blocks."""
    from i3rc_monte_carlo_model_b200._lib import backend
    from tests.philox_numpy import photon_block, u01
    be = backend()
    ph = np.array([0], np.uint64)
    bl = np.array([0], np.uint32)
    raw, u = np.zeros(4, np.uint32), np.zeros(4, np.float32)
    U64, U32 = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    assert be.probe_philox(0, 0, 1, ph.ctypes.data_as(U64), bl.ctypes.data_as(U32), raw.ctypes.data_as(U32), _abi.fptr(u)) == 0
    assert raw.tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    rng = np.random.default_rng(11)
    n = 4096
    ph = rng.integers(0, 2**40, n).astype(np.uint64)
    bl = rng.integers(0, 5000, n).astype(np.uint32)
    for key in ((10, 1), (0xFFFFFFFF, 0x9E3779B9), (123456789, 80)):
        raw, u = np.zeros((n, 4), np.uint32), np.zeros((n, 4), np.float32)
        assert be.probe_philox(key[0], key[1], n, ph.ctypes.data_as(U64), bl.ctypes.data_as(U32), raw.ctypes.data_as(U32),
                               _abi.fptr(u)) == 0
        want = photon_block(key, ph, bl)
        assert np.array_equal(raw, want)
        assert np.array_equal(u, u01(want))  # the deviates: bit-identical float32
        assert u.min() >= 0.0 and u.max() <= 1.0


def test_next_direct_on_the_device_matches_the_oracle(oracle):
    """next_direct (MCRT:2086-2113) exactly as the kernel runs it -- rejection rounds fed by the photon's Philox blocks --
    against the oracle's next_direct fed with the same deviates (numpy Philox): 1e-6 absolute on direction cosines."""
    from i3rc_monte_carlo_model_b200._lib import backend
    from tests.philox_numpy import photon_block, u01
    be = backend()
    rng = np.random.default_rng(3)
    n = 4000
    S = rng.standard_normal((n, 3))
    S[:5] = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0], [0.6, 0, 0.8]]
    S = (S / np.linalg.norm(S, axis=1, keepdims=True)).astype(np.float32)
    cs = np.cos(rng.random(n) * np.pi).astype(np.float32)
    cs[:3] = [1.0, -1.0, 0.0]
    ph = rng.integers(0, 2**33, n).astype(np.uint64)
    bl = rng.integers(0, 100, n).astype(np.uint32)
    key = (10, 7)
    out, used = np.zeros((n, 3), np.float32), np.zeros(n, np.uint32)
    U64, U32 = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    assert be.probe_next_direct(key[0], key[1], n, ph.ctypes.data_as(U64), bl.ctypes.data_as(U32), _abi.fptr(S), _abi.fptr(cs),
                                _abi.fptr(out), used.ctypes.data_as(U32)) == 0
    nblk = 8  # deviates of 8 blocks = 16 rejection rounds: (1 - pi/4)^16 ~ 2e-11 that it is not enough
    xi = np.concatenate([u01(photon_block(key, ph, bl.astype(np.uint64) + k)) for k in range(nblk)], axis=1)  # [n, 4*nblk]
    want = S.copy()
    rounds = np.zeros(n, int)
    for i in range(n):
        x = np.ascontiguousarray(xi[i])
        oracle.lib.orc_next_direct(_abi.fptr(x), x.size, float(cs[i]), _abi.fptr(want[i]))
        ax, ay = 1 - 2 * x[0::2], 1 - 2 * x[1::2]
        rounds[i] = np.argmax(ax * ax + ay * ay <= 1.0) + 1
    assert np.array_equal(used, (rounds + 1) // 2)  # two rounds per Philox block
    assert np.max(np.abs(out - want)) < 2e-6
    assert np.max(np.abs(np.linalg.norm(out.astype(np.float64), axis=1) - 1.0)) < 1e-5
    dots = np.sum(out.astype(np.float64) * S, axis=1)
    assert np.max(np.abs(dots - cs)) < 1e-5  # the new direction makes the scattering angle with the old one


def test_components_are_validated_inside_the_c_abi(cuda):
    """validateOpticalComponent's value checks (Code/opticalProperties.f95:966-975) hold for callers that hand file data
    straight to the C ABI (the C++ drivers, the Fortran shim): the arrays are altered behind the Python mirror's back."""
    for field, value, text in (("extinction", -1.0, "extinction must be >= 0"),
                               ("singleScatteringAlbedo", 1.5, "singleScatteringAlbedo must be between 0 and 1"),
                               ("phaseFunctionIndex", 2, "phase function index is out of bounds"),
                               ("phaseFunctionIndex", -1, "phase function index is out of bounds")):
        d = fields.step_cloud(0.99)
        getattr(d.components[0], field)[3, 0, 5] = value
        st = ErrorMessage()
        I = new_Integrator(d, status=st, backend=cuda)
        assert stateIsFailure(st) and not I.handle
        assert text in cuda.last_message(None).decode(), (field, cuda.last_message(None))


def test_tallies_staged_in_shared_memory_equal_global_atomics(cuda):
    """Few-column domains keep per-warp tallies in shared memory (warp-aggregated, then one block-level sum and one
    global atomic per element); the result must equal the plain global-atomic path photon for photon (same Philox
    streams), up to float32 summation order."""
    for make, kw in ((lambda: fields.plane_parallel(), dict(useRussianRouletteForIntensity=False)),
                     (lambda: fields.step_cloud(0.99), dict(useRussianRouletteForIntensity=True, zetaMin=0.3)),
                     (lambda: fields.plane_parallel(nX=3, nY=2, nLayers=4, SSA=0.9), dict())):
        res = []
        for stage in (1536, 0):
            I = make_integrator(cuda, make(), surfaceAlbedo=0.3, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0], **kw)
            assert cuda.set_tuning(I.handle, b"stage_tallies", stage) == 0
            assert cuda.set_tuning(I.handle, b"stage_max_columns", 64) == 0  # (the step cloud's 32 columns too)
            assert (cuda.get_layout(I.handle, 1) > 0) == (stage > 0)
            ph = new_PhotonStream(0.5, 0.0, numberOfPhotons=60_000)
            computeRadiativeTransfer(I, new_RandomNumberSequence([10, 1]), ph)
            r = reportResults(I, "fluxUp", "fluxDown", "fluxAbsorbed", "intensity", "volumeAbsorption", "meanFluxUp")
            res.append((r, getCounters(I)))
        (a, ca), (b, cb) = res
        assert ca == cb
        # (a single float32 element that takes 1e5 .. 1e6 increments by global atomics has itself lost up to 1e-3 of its
        #  value -- measured: 1.1e-3 on the one-column radiance at 3e5 photons; the staged sums, many short ones, are
        #  the more accurate)
        for k in ("fluxUp", "fluxDown", "fluxAbsorbed", "intensity", "volumeAbsorption"):
            assert np.allclose(a[k], b[k], rtol=2e-3, atol=1e-7), (k, np.max(np.abs(a[k] - b[k])))


def test_births_in_groups_trace_the_same_photons(cuda):
    """A warp starts new photons when `birth_min` of its slots are empty (k_transport, kernels.cuh) instead of refilling
    every slot at once: which slot and which lane a photon gets changes, its Philox stream does not, so every counter is
    the same and the results agree to float32 summation order -- on a small domain (staged tallies), on a domain of a
    few columns and on the Landsat field, with a photon count that does not fill the last group."""
    kw = dict(surfaceAlbedo=0.2, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0], useRussianRouletteForIntensity=True,
              zetaMin=0.3)
    for make, nph in ((lambda: fields.plane_parallel(nX=3, nY=2, nLayers=4, SSA=0.9), 70_001),
                      (lambda: fields.step_cloud(0.99), 100_003), (lambda: fields.landsat_cloud(0.98, nLegendreCoefficients=16), 300_007)):
        res = []
        for bmin, blow in ((1, 16), (16, 4), (32, 0)):
            I = make_integrator(cuda, make(), **kw)
            assert cuda.set_tuning(I.handle, b"birth_min", bmin) == 0 and cuda.set_tuning(I.handle, b"birth_low", blow) == 0
            computeRadiativeTransfer(I, new_RandomNumberSequence([7, 2]), new_PhotonStream(0.5, 30.0, numberOfPhotons=nph))
            res.append((reportResults(I, "fluxUp", "fluxDown", "fluxAbsorbed", "intensity", "volumeAbsorption"), getCounters(I)))
        (a, ca) = res[0]
        assert ca["photons"] == nph and ca["bad"] == 0
        for b, cb in res[1:]:
            assert ca == cb
            for k in a:
                assert np.allclose(a[k], b[k], rtol=2e-3, atol=1e-6), (k, np.max(np.abs(a[k] - b[k])))


def test_eighty_slots_per_warp_trace_the_same_photons(cuda):
    """Domains of many columns run a kernel with 80 photon slots per warp whose scratch of suspended event batches lives in
    global memory (k_transport, SuspScratch): same photons, same Philox streams as with 64 slots -- on the Landsat field
    and on a two-component field with four radiance directions (the kernel is chosen for domains of at least 4096 columns
    and at most four directions)."""
    cases = ((lambda: fields.landsat_cloud(0.98, nLegendreCoefficients=16),
              dict(intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0]), 300_007),
             (lambda: fields.synthetic_les(nx=64, ny=64, nz=32),
              dict(intensityMus=[1.0, 0.8, 0.6, 0.4], intensityPhis=[0.0, 90.0, 180.0, 270.0]), 100_003))
    for make, dirs, nph in cases:
        res = []
        for s80 in (1, 0):
            I = make_integrator(cuda, make(), surfaceAlbedo=0.2, useRussianRouletteForIntensity=True, zetaMin=0.3, **dirs)
            assert cuda.set_tuning(I.handle, b"slots_80", s80) == 0
            computeRadiativeTransfer(I, new_RandomNumberSequence([7, 3]), new_PhotonStream(0.5, 30.0, numberOfPhotons=nph))
            res.append((reportResults(I, "fluxUp", "fluxDown", "fluxAbsorbed", "intensity", "volumeAbsorption"), getCounters(I)))
        (a, ca), (b, cb) = res
        assert ca["photons"] == nph and ca["bad"] == 0 and ca == cb
        for k in a:
            assert np.allclose(a[k], b[k], rtol=2e-3, atol=1e-6), (k, np.max(np.abs(a[k] - b[k])))


def test_empty_space_codes_change_nothing_but_the_number_of_gathers(cuda, monkeypatch):
    """Rays that jump through empty space (the coded copy of the gather field, transport.cuh JUMP_*) trace the same
    photons as rays that look at every cell: same Philox streams, so the batch results agree to float32 rounding (a jump
    re-derives the distances to the next faces, which can move an exit point by one ulp), the event counters agree, and
    DDA steps + cells skipped equals the number of steps of the plain kernel."""
    monkeypatch.setenv("I3RC_SKIP_EMPTY", "1")
    kw = dict(surfaceAlbedo=0.1, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0], useRussianRouletteForIntensity=True,
              zetaMin=0.3)
    d = fields.landsat_cloud(1.0, nLegendreCoefficients=16)
    res = []
    for skip in (1, 0):
        I = make_integrator(cuda, d, **kw)
        assert cuda.get_layout(I.handle, 2) > 20, "the Landsat field has empty space to code"
        if not skip:
            assert cuda.set_tuning(I.handle, b"skip_empty", 0) == 0 and cuda.get_layout(I.handle, 2) == 0
        computeRadiativeTransfer(I, new_RandomNumberSequence([10, 3]), new_PhotonStream(0.5, 0.0, numberOfPhotons=400_000))
        res.append((reportResults(I, "meanFluxUp", "meanFluxDown", "meanIntensity", "fluxUp", "intensity"), getCounters(I)))
    (a, ca), (b, cb) = res
    assert ca["bad"] == 0 and cb["cells_skipped"] == 0 and ca["cells_skipped"] + ca["cells_skipped_intensity"] > 0.05 * cb["crossings_photon"]
    steps = lambda c: c["crossings_photon"] + c["crossings_intensity"] + c["cells_skipped"] + c["cells_skipped_intensity"]
    assert abs(steps(ca) - steps(cb)) <= 2e-4 * steps(cb)
    for k in ("collisions", "exits_top", "surface_hits", "contributions"):
        assert abs(ca[k] - cb[k]) <= 2e-4 * cb[k] + 2, k
    for k in ("meanFluxUp", "meanFluxDown", "meanIntensity"):
        assert np.allclose(a[k], b[k], rtol=2e-4), k
    assert np.abs(a["fluxUp"] - b["fluxUp"]).mean() < 0.01 * b["fluxUp"].mean()


def test_straight_up_directions_from_column_sums_equal_traced_rays_on_the_device(cuda):
    """The nadir view (mu = 1) integrated from column suffix sums instead of traced: same photons, same contributions up
    to float32 summation order, fewer cell crossings -- on a two-component field with Russian roulette for intensity and
    on an irregular grid with the plain local estimate."""
    cases = [(fields.synthetic_les(nx=24, ny=16, nz=32, n_entries=3, seed=7, nLegendreCoefficients=16),
              dict(surfaceAlbedo=0.1, intensityMus=[1.0, 0.6, 1.0], intensityPhis=[0.0, 40.0, 180.0], useRussianRouletteForIntensity=True,
                   zetaMin=0.3)),
             (_irregular_domain(), dict(surfaceAlbedo=0.3, intensityMus=[0.7, 1.0], intensityPhis=[10.0, 0.0],
                                        useRussianRouletteForIntensity=False))]
    for d, kw in cases:
        res = []
        for v in (1, 0):
            I = make_integrator(cuda, d, **kw)
            assert cuda.set_tuning(I.handle, b"vertical_shortcut", v) == 0
            computeRadiativeTransfer(I, new_RandomNumberSequence([10, 2]), new_PhotonStream(0.6, 20.0, numberOfPhotons=200_000))
            res.append((reportResults(I, "intensity", "fluxUp", "meanIntensity"), getCounters(I)))
        (a, ca), (b, cb) = res
        assert np.allclose(a["meanIntensity"], b["meanIntensity"], rtol=1e-4)
        assert np.allclose(a["intensity"], b["intensity"], rtol=2e-3, atol=1e-6)
        assert np.allclose(a["fluxUp"], b["fluxUp"], rtol=1e-4, atol=1e-7)
        for k in ("crossings_photon", "collisions", "contributions", "rng_draws", "exits_top", "surface_hits"):
            assert ca[k] == cb[k], k
        assert ca["crossings_intensity"] < 0.9 * cb["crossings_intensity"]


def test_large_hand_filled_photon_arrays_are_traced_in_overlapped_pieces(cuda):
    """More than 2^20 hand-filled photons: the arrays are uploaded piece by piece on a copy stream while the piece before
    is traced.  Every photon is traced exactly once (counter), and the result agrees with the device-drawn source."""
    from tests.cases import familywise_bound
    d = fields.step_cloud(0.99)
    n, nb = 1_300_000, 6
    kw = dict(surfaceAlbedo=0.1, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0], useRussianRouletteForIntensity=True, zetaMin=0.3)
    res = {}
    for how in ("arrays", "descriptor"):
        I = make_integrator(cuda, d, **kw)
        vals = []
        for b in range(nb):
            ph = new_PhotonStream(0.5, 0.0, numberOfPhotons=n)
            if how == "arrays":
                r2 = np.random.default_rng(500 + b)
                ph.xPosition, ph.yPosition = r2.random(n, dtype=np.float32), r2.random(n, dtype=np.float32)
                ph.zPosition = np.full(n, 1.0 - np.finfo(np.float32).eps, np.float32)
                ph.initialMu, ph.initialPhi = np.full(n, -0.5, np.float32), np.zeros(n, np.float32)
            computeRadiativeTransfer(I, new_RandomNumberSequence([10, b + 1]), ph)
            c = getCounters(I)
            assert c["photons"] == n and c["bad"] == 0
            r = reportResults(I, "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity")
            vals.append([r["meanFluxUp"], r["meanFluxDown"], r["meanFluxAbsorbed"], *r["meanIntensity"]])
        res[how] = np.array(vals, np.float64)
    a, b = res["arrays"], res["descriptor"]
    z = (a.mean(0) - b.mean(0)) / np.sqrt(a.var(0, ddof=1) / nb + b.var(0, ddof=1) / nb)
    assert np.all(np.abs(z) <= familywise_bound(z.size, nb - 1)), z
    closure = a[:, 0] + a[:, 2] + 0.9 * a[:, 1]
    assert abs(closure.mean() - 1.0) < 2e-3


def test_phase_tables_staged_in_shared_memory_give_the_same_photons(cuda):
    """The experimental one-block-per-SM kernel that stages the inverse and forward phase-function tables in shared memory
    (`tables_in_smem`): same Philox streams, same table values, so the same batch up to float32 summation order."""
    d = fields.landsat_cloud(1.0, nLegendreCoefficients=32)
    kw = dict(surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0], useRussianRouletteForIntensity=True,
              zetaMin=0.3, minInverseTableSize=10001, minForwardTableSize=10001)
    res = []
    for staged in (1, 0):
        I = make_integrator(cuda, d, **kw)
        assert cuda.set_tuning(I.handle, b"tables_in_smem", staged) == 0
        computeRadiativeTransfer(I, new_RandomNumberSequence([10, 4]), new_PhotonStream(0.5, 0.0, numberOfPhotons=300_000))
        res.append((reportResults(I, "meanFluxUp", "meanFluxDown", "meanIntensity", "fluxUp"), getCounters(I)))
    (a, ca), (b, cb) = res
    assert ca == cb
    for k in ("meanFluxUp", "meanFluxDown", "meanIntensity"):
        assert np.allclose(a[k], b[k], rtol=1e-5), k
    assert np.allclose(a["fluxUp"], b["fluxUp"], rtol=1e-4, atol=1e-6)


def test_uniform_slabs_crossed_in_one_go_change_nothing_but_the_number_of_steps(cuda, force_layer_split):
    """Runs of horizontally uniform layers (clear air with a gas component above and below the clouds) are crossed in one
    step when the ray's optical-path limit lies beyond them (transport.cuh, ray_cross_slab).  Against the same kernel
    walking them cell by cell (`slab_jump` = 0): fixed rays end in the same cell with the same optical path (1e-5), a
    photon batch gives the same tallies up to float32 rounding, and steps + cells skipped is conserved."""
    d = fields.synthetic_les(nx=24, ny=16, nz=64, n_entries=3, seed=7, nLegendreCoefficients=16)
    kw = dict(surfaceAlbedo=0.2, intensityMus=[1.0, 0.6, 0.3, 0.8], intensityPhis=[0.0, 40.0, 200.0, 310.0],
              useRussianRouletteForIntensity=True, zetaMin=0.3)
    rng = np.random.default_rng(21)
    n = 2000
    hi = np.array([d.xPosition[-1], d.yPosition[-1], d.zPosition[-1]], np.float64)
    pos = ((0.01 + 0.98 * rng.random((n, 3))) * hi).astype(np.float32)
    mu = rng.uniform(0.05, 1.0, n) * rng.choice([-1, 1], n)
    phi = rng.uniform(0, 2 * np.pi, n)
    u = np.column_stack([np.sqrt(1 - mu**2) * np.cos(phi), np.sqrt(1 - mu**2) * np.sin(phi), mu]).astype(np.float32)
    u[:4] = [[0, 0, 1], [0, 0, -1], [0.6, 0, 0.8], [0, -0.8, -0.6]]
    lim = np.where(rng.random(n) < 0.5, np.inf, rng.exponential(2.0, n)).astype(np.float32)
    res = []
    for jump in (1, 0):
        I = make_integrator(cuda, d, **kw)
        assert cuda.get_layout(I.handle, 0) > 0, "the layer table is in use"
        assert cuda.set_tuning(I.handle, b"slab_jump", jump) == 0
        rays = traceRays(I, pos, u, lim)
        computeRadiativeTransfer(I, new_RandomNumberSequence([10, 5]), new_PhotonStream(0.5, 30.0, numberOfPhotons=300_000))
        res.append((rays, reportResults(I, "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity", "fluxUp", "intensity"),
                    getCounters(I)))
    (ra, a, ca), (rb, b, cb) = res
    assert np.max(np.abs(ra[0] - rb[0]) / np.maximum(np.abs(rb[0]), 1e-3)) < 1e-5           # optical paths
    assert np.mean(np.all(ra[2] == rb[2], axis=1)) > 0.995 and np.array_equal(ra[2][:, 2], rb[2][:, 2])  # end cells (z: always)
    assert cb["cells_skipped"] == 0 and cb["cells_skipped_intensity"] == 0 and ca["bad"] == 0 and cb["bad"] == 0
    assert ca["cells_skipped"] + ca["cells_skipped_intensity"] > 0.2 * (cb["crossings_photon"] + cb["crossings_intensity"])
    tot = lambda c: c["crossings_photon"] + c["crossings_intensity"] + c["cells_skipped"] + c["cells_skipped_intensity"]
    assert abs(tot(ca) - tot(cb)) <= 2e-3 * tot(cb)
    for k in ("collisions", "exits_top", "surface_hits", "contributions", "absorptions"):
        assert abs(ca[k] - cb[k]) <= 3e-4 * cb[k] + 2, k
    for k in ("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity"):
        assert np.allclose(a[k], b[k], rtol=3e-4), k


def test_rays_that_cannot_contribute_are_not_traced_on_the_device(cuda):
    """The lower bound of the optical path to the top (Problem::leLB) on the device: with and without it the same photons
    give the same tallies (float32 summation order aside) and the same number of contributions; only the local-estimate
    crossings fall.  Landsat cloud with the bench.py parameters, and a two-component field with five directions (one of
    them downward: under roulette it can never contribute, quirk Q4)."""
    cases = [(fields.landsat_cloud(1.0, nLegendreCoefficients=32),
              dict(surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0], useRussianRouletteForIntensity=True,
                   zetaMin=0.3), dict(solarMu=0.5, solarAzimuth=0.0)),
             (fields.synthetic_les(nx=24, ny=16, nz=32, n_entries=3, seed=7, nLegendreCoefficients=16),
              dict(surfaceAlbedo=0.2, intensityMus=[1.0, 0.5, 0.5, 0.8, -0.6], intensityPhis=[0.0, 0.0, 180.0, 130.0, 20.0],
                   useRussianRouletteForIntensity=True, zetaMin=0.3), dict(solarMu=0.5, solarAzimuth=30.0))]
    for d, kw, src in cases:
        res = []
        for lb in (1, 0):
            I = make_integrator(cuda, d, **kw)
            assert cuda.set_tuning(I.handle, b"le_lower_bound", 2 * lb) == 0  # (2: also on domains of few columns)
            computeRadiativeTransfer(I, new_RandomNumberSequence([10, 6]), new_PhotonStream(numberOfPhotons=400_000, **src))
            res.append((reportResults(I, "meanIntensity", "intensity", "fluxUp", "meanFluxUp"), getCounters(I)))
        (a, ca), (b, cb) = res
        for k in ("crossings_photon", "collisions", "contributions", "rng_draws", "exits_top", "surface_hits"):
            assert ca[k] == cb[k], k
        assert ca["crossings_intensity"] < 0.85 * cb["crossings_intensity"]
        assert np.allclose(a["meanIntensity"], b["meanIntensity"], rtol=1e-5)
        # (a ray that is certain to survive the roulette is tallied where the straight line leaves the domain: the traced
        #  ray may leave one column further when that point lies within an ulp of a column boundary)
        assert np.mean(~np.isclose(a["intensity"], b["intensity"], rtol=2e-3, atol=1e-6)) < 1e-3
        assert np.array_equal(a["fluxUp"] > 0, b["fluxUp"] > 0) and np.allclose(a["fluxUp"], b["fluxUp"], rtol=1e-4, atol=1e-7)
        # with the upper bound as well (off by default): fewer crossings again, same contributions
        I = make_integrator(cuda, d, **kw)
        assert cuda.set_tuning(I.handle, b"le_lower_bound", 2) == 0 and cuda.set_tuning(I.handle, b"le_upper_bound", 1) == 0
        computeRadiativeTransfer(I, new_RandomNumberSequence([10, 6]), new_PhotonStream(numberOfPhotons=400_000, **src))
        cc, rc = getCounters(I), reportResults(I, "meanIntensity", "intensity")
        assert cc["contributions"] == ca["contributions"] and cc["crossings_intensity"] < ca["crossings_intensity"]
        assert np.allclose(rc["meanIntensity"], b["meanIntensity"], rtol=1e-5)
        assert np.mean(~np.isclose(rc["intensity"], b["intensity"], rtol=2e-3, atol=1e-6)) < 1e-3
