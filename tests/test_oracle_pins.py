"""Pins of the CPU oracle.  The reference ships no golden vectors and cannot be compiled here (PARITY UNPINNED,
see oracle/i3rc_oracle.h), so the oracle is pinned to (i) the public MT19937 known-answer vectors and numpy's
bit-identical generator, (ii) independent numerical libraries for the deterministic sub-paths, and (iii)
analytic known answers of radiative transfer."""
import ctypes as C

import numpy as np
import pytest
from numpy.polynomial import legendre as npleg

from i3rc_monte_carlo_model_b200 import _abi, fields
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, getTable, new_Integrator,
                                                                    reportResults, specifyParameters, traceRays)
from i3rc_monte_carlo_model_b200.opticalProperties import addOpticalComponent, new_Domain
from i3rc_monte_carlo_model_b200.RandomNumbers import new_RandomNumberSequence
from i3rc_monte_carlo_model_b200.scatteringPhaseFunctions import new_PhaseFunction, new_PhaseFunctionTable
from oracle.binding import MT19937
from tests.cases import make_integrator, mean_se, run_batches


def test_mt19937_public_known_answer():
    # mt19937ar.c reference output for init_by_array({0x123, 0x234, 0x345, 0x456})
    m = MT19937([0x123, 0x234, 0x345, 0x456])
    assert [m.int32() for _ in range(5)] == [1067595299, 955945823, 477289528, 4107218783, 4228976476]


@pytest.mark.parametrize("seed", [[10, 1], [10, 0], [3, 10], [12345, 678]])
def test_mt19937_matches_numpy_legacy_seeding(seed):
    """RandomNumbersForMC.f95:187-299 == numpy's MT19937 with init_by_array; getRandomReal = genrand_real1 -> float32."""
    m = MT19937(seed)
    bg = np.random.MT19937()
    bg._legacy_seeding(np.array(seed, dtype=np.uint32))
    raw = bg.random_raw(2000)
    mine = np.array([m.real() for _ in range(2000)], dtype=np.float32)
    expect = (raw.astype(np.float64) / (2.0**32 - 1.0)).astype(np.float32)
    assert np.array_equal(mine, expect)


def test_mt19937_scalar_seed():
    m = MT19937(5489)
    assert m.int32() == 3499211612  # mt19937 default-seed first output


def test_findIndex_semantics(oracle):
    rng = np.random.default_rng(0)
    table = np.sort(rng.random(50)).astype(np.float32)
    for v in np.concatenate([rng.random(200).astype(np.float32), table[:5]]):
        got = oracle.lib.orc_findIndex(float(v), _abi.fptr(table), table.size, 0)
        expect = int(np.searchsorted(table, v, side="right"))  # 1-based index of last entry <= v; 0 if below
        assert got == expect
        guess = int(rng.integers(1, table.size))  # "the firstGuess always makes sense" (numericUtilities.f95:204-205)
        if expect >= 1:
            assert oracle.lib.orc_findIndex(float(v), _abi.fptr(table), table.size, guess) == expect


@pytest.mark.parametrize("n", [2, 3, 8, 64, 299])
def test_lobatto_nodes(oracle, n):
    """numericUtilities.f95:15-102: nodes are -1, +1 and the roots of P'_{n-1}."""
    mus = np.zeros(n, np.float32)
    oracle.lib.orc_computeLobattoMus(_abi.fptr(mus), n)
    c = np.zeros(n)
    c[-1] = 1.0
    roots = np.sort(npleg.legroots(npleg.legder(c))) if n > 2 else np.array([])
    expect = np.concatenate([[-1.0], roots, [1.0]])
    assert np.allclose(mus, expect, atol=3e-6)


def test_legendre_recursion(oracle):
    mus = np.linspace(-1, 1, 41).astype(np.float32)
    maxL = 64
    P = np.zeros((mus.size, maxL + 1), np.float32)
    oracle.lib.orc_computeLegendrePolynomials.argtypes = [C.c_int, _abi.c_float_p, C.c_int, _abi.c_float_p]
    oracle.lib.orc_computeLegendrePolynomials(maxL, _abi.fptr(mus), mus.size, _abi.fptr(P))
    for l in (0, 1, 2, 7, 33, 64):
        c = np.zeros(l + 1)
        c[-1] = 1
        assert np.allclose(P[:, l], npleg.legval(mus.astype(np.float64), c), atol=2e-5)


def _hg_mu_of_cdf(g, p):
    """Closed-form inverse of the Henyey-Greenstein CDF measured from mu = -1."""
    return (1 + g * g - ((1 - g * g) / (1 - g + 2 * g * p)) ** 2) / (2 * g)


def test_inverse_table_matches_HG_closed_form(oracle):
    g = 0.85
    I = make_integrator(oracle, fields.plane_parallel(g=g, nLegendreCoefficients=64), surfaceAlbedo=0.0)
    oracle.tabulate(I.handle)
    T = getTable(I, 0)[0]
    n = T.size
    assert n == 9001 and T[0] == pytest.approx(np.pi, abs=2e-3) and T[-1] == 0.0
    p = np.arange(n) / (n - 1)
    mu = _hg_mu_of_cdf(g, p)
    # the table inverts a 64-point Lobatto/trapezoid CDF of a 64-moment series: ~1e-2 max, ~1e-3 mean error in mu
    sel = slice(1, n - 1)
    err = np.abs(np.cos(T[sel]) - mu[sel])
    assert err.max() < 1.5e-2 and err.mean() < 2e-3
    assert np.all(np.diff(T) <= 1e-6)  # monotone non-increasing


def test_forward_table_matches_HG(oracle):
    g = 0.85
    I = make_integrator(oracle, fields.plane_parallel(g=g, nLegendreCoefficients=299), surfaceAlbedo=0.0,
                        intensityMus=[1.0], intensityPhis=[0.0])
    oracle.tabulate(I.handle)
    F = getTable(I, 1)[0].astype(np.float64)
    th = np.linspace(0, np.pi, F.size)
    hg = (1 - g * g) / (1 + g * g - 2 * g * np.cos(th)) ** 1.5
    assert np.max(np.abs(F - hg) / hg) < 2e-3  # float32 Legendre sum of 299 terms
    # normalised to integral P dmu = 2
    assert np.trapezoid(F * np.sin(th), th) == pytest.approx(2.0, rel=1e-4)


def test_tabulated_phase_function_is_normalised(oracle):
    I = make_integrator(oracle, fields.plane_parallel(useMoments=False, nAngles=5000), surfaceAlbedo=0.0,
                        intensityMus=[1.0], intensityPhis=[0.0])
    oracle.tabulate(I.handle)
    F = getTable(I, 2)[0].astype(np.float64)
    th = np.linspace(0, np.pi, F.size)
    assert np.trapezoid(F * np.sin(th), th) == pytest.approx(2.0, rel=2e-4)


def test_hybrid_phase_function(oracle):
    """computeHydridPhaseFunctions (MCRT:1925-1998): Gaussian forward peak, continuous, still normalised to 2."""
    # the Deirmendjian C1 function has a sharp enough peak for a 7 degree Gaussian to take over (for HG, g = 0.85,
    # the search of MCRT:1961-1975 finds no transition and leaves the table untouched)
    I = make_integrator(oracle, fields.radar_cloud(1.0, "C1"), surfaceAlbedo=0.0, intensityMus=[1.0],
                        intensityPhis=[0.0], useHybridPhaseFunsForIntenCalcs=True, hybridPhaseFunWidth=7.0)
    oracle.tabulate(I.handle)
    H, O = getTable(I, 1)[0].astype(np.float64), getTable(I, 2)[0].astype(np.float64)
    th = np.linspace(0, np.pi, H.size)
    k = np.nonzero(H != O)[0]
    assert k.size > 10 and k.max() < H.size // 3  # only the forward peak is replaced
    assert H[0] < O[0]  # peak is smoothed
    assert np.trapezoid(H * np.sin(th), th) == pytest.approx(2.0, rel=2e-3)
    t = k.max()
    assert abs(H[t] - O[t + 1]) / O[t + 1] < 0.02  # continuous at the transition


def _f64_optical_path(d, tot, pos, direction, tmax=None):
    """Exact (float64) optical path of a ray through the gridded extinction to the top/bottom boundary."""
    xe, ye, ze = (np.asarray(a, np.float64) for a in (d.xPosition, d.yPosition, d.zPosition))
    nx, ny, nz = xe.size - 1, ye.size - 1, ze.size - 1
    p = np.array(pos, np.float64)
    u = np.array(direction, np.float64)
    Lx, Ly = xe[-1] - xe[0], ye[-1] - ye[0]
    tau = 0.0
    ix = min(np.searchsorted(xe, p[0], side="right") - 1, nx - 1)
    iy = min(np.searchsorted(ye, p[1], side="right") - 1, ny - 1)
    iz = min(np.searchsorted(ze, p[2], side="right") - 1, nz - 1)
    for _ in range(200000):
        ts = []
        for a, (e, i) in enumerate(((xe, ix), (ye, iy), (ze, iz))):
            if abs(u[a]) < 1e-300:
                ts.append(np.inf)
            else:
                ts.append(((e[i + 1] if u[a] > 0 else e[i]) - p[a]) / u[a])
        t = max(min(ts), 0.0)
        tau += t * tot[ix, iy, iz]
        p = p + t * u
        a = int(np.argmin(ts))
        if a == 0:
            ix += 1 if u[0] > 0 else -1
            if ix >= nx:
                ix, p[0] = 0, p[0] - Lx
            elif ix < 0:
                ix, p[0] = nx - 1, p[0] + Lx
        elif a == 1:
            iy += 1 if u[1] > 0 else -1
            if iy >= ny:
                iy, p[1] = 0, p[1] - Ly
            elif iy < 0:
                iy, p[1] = ny - 1, p[1] + Ly
        else:
            iz += 1 if u[2] > 0 else -1
            if iz >= nz or iz < 0:
                return tau
    raise RuntimeError("ray did not leave the domain")


def test_fixed_ray_optical_path_vs_float64(oracle):
    """accumulateExtinctionAlongPath (MCRT:1654-1807) against an exact float64 integration (8c-vi)."""
    from tests.hostsim.binding import dense_from_domain
    d = fields.step_cloud(1.0)
    tot = dense_from_domain(d)[0].astype(np.float64)
    I = make_integrator(oracle, d, surfaceAlbedo=0.0)
    rng = np.random.default_rng(5)
    n = 200
    pos = np.column_stack([rng.uniform(1, 499, n), rng.uniform(1, 499, n), rng.uniform(1, 249, n)]).astype(np.float32)
    mu = rng.uniform(0.2, 1.0, n) * rng.choice([-1, 1], n)
    phi = rng.uniform(0, 2 * np.pi, n)
    u = np.column_stack([np.sqrt(1 - mu**2) * np.cos(phi), np.sqrt(1 - mu**2) * np.sin(phi), mu]).astype(np.float32)
    tau, _, _ = traceRays(I, pos, u)
    exact = np.array([_f64_optical_path(d, tot, pos[i], u[i]) for i in range(n)])
    assert np.all(tau >= 0)
    assert np.max(np.abs(tau - exact) / exact) < 2e-5


# ---- analytic radiative-transfer answers ---------------------------------------------------------------------
def _slab(oracle, tau=1.0, ssa=1.0, g=0.0, nmom=2, **params):
    coefs = [g**l for l in range(1, nmom + 1)]
    table = new_PhaseFunctionTable([new_PhaseFunction(np.array(coefs, np.float32))], [1.0])
    d = new_Domain([0.0, 500.0], [0.0, 500.0], [0.0, 250.0])
    ext = np.full((1, 1, 1), tau / 250.0, np.float32)
    addOpticalComponent(d, "slab", ext, np.full_like(ext, ssa), np.ones((1, 1, 1), np.int32), table)
    return make_integrator(oracle, d, **params)


def test_beer_law_pure_absorber(oracle):
    """ssa = 0: fluxDown = exp(-tau/mu0), fluxUp = 0 (8c-iv); absorption closes the budget."""
    I = _slab(oracle, tau=1.5, ssa=0.0, surfaceAlbedo=0.0, useRussianRoulette=False)
    r = run_batches(I, 50000, 8, source=dict(solarMu=0.6, solarAzimuth=30.0))
    m, s = mean_se(r["meanFluxDown"])
    assert abs(m - np.exp(-1.5 / 0.6)) < 4 * s + 1e-6
    assert np.all(r["meanFluxUp"] == 0)
    assert np.allclose(r["meanFluxAbsorbed"] + r["meanFluxDown"], 1.0, atol=1e-5)


def test_energy_closure_and_absorption_consistency(oracle):
    """fluxUp + fluxAbsorbed + (1-albedo) fluxDown = 1 in expectation (8c-ii);
    fluxAbsorbed(x,y) = sum_k volumeAbsorption(x,y,k) dz_k to round-off (8c-iii)."""
    d = fields.step_cloud(0.99)
    I = make_integrator(oracle, d, surfaceAlbedo=0.3)
    r = run_batches(I, 40000, 8, want=["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxAbsorbed", "volumeAbsorption"])
    closure = r["meanFluxUp"] + r["meanFluxAbsorbed"] + 0.7 * r["meanFluxDown"]
    m, s = mean_se(closure)
    assert abs(m - 1.0) < 4 * s + 1e-4  # Russian roulette makes it hold in expectation only
    dz = np.diff(d.zPosition)
    col = (r["volumeAbsorption"] * dz[None, None, None, :]).sum(-1)
    assert np.allclose(col, r["fluxAbsorbed"], rtol=2e-4, atol=1e-6)


def test_lambertian_surface_under_transparent_atmosphere(oracle):
    """I = albedo * solarFlux / pi for every upward direction, no mu0 factor (8c-viii, quirk Q10)."""
    I = _slab(oracle, tau=1e-6, ssa=1.0, surfaceAlbedo=0.4, intensityMus=[1.0, 0.5, 0.2], intensityPhis=[0.0, 90.0, 270.0])
    r = run_batches(I, 20000, 4, source=dict(solarMu=0.5, solarAzimuth=0.0))
    m, _ = mean_se(r["meanIntensity"])
    assert np.allclose(m, 0.4 / np.pi, rtol=2e-4)


def test_single_scatter_isotropic_slab_radiance(oracle):
    """Thin isotropic slab: nadir-view reflected radiance ~ single scattering closed form (8c-vii):
    I(mu) = w0 P/(4 pi) * mu0/(mu0+mu) * (1 - exp(-tau (1/mu0 + 1/mu))) / mu0 * mu0 ... in flux-normalised units."""
    tau, w0, mu0, mu = 0.05, 1.0, 0.5, 1.0
    I = _slab(oracle, tau=tau, ssa=w0, g=0.0, nmom=2, surfaceAlbedo=0.0, intensityMus=[mu], intensityPhis=[0.0],
              useRussianRouletteForIntensity=False)
    r = run_batches(I, 100000, 8, source=dict(solarMu=mu0, solarAzimuth=0.0))
    m, s = mean_se(r["meanIntensity"])
    # photons carry unit weight per horizontal area (flux on the horizontal = 1): F0 = 1/mu0; P = 1 (normalised to 2 -> isotropic P = 1)
    single = w0 * 1.0 / (4 * np.pi) * (1.0 / mu0) * (mu0 / (mu0 + mu)) * (1 - np.exp(-tau * (1 / mu0 + 1 / mu)))
    # multiple scattering adds O(tau) relative: allow 12 %
    assert m[0] == pytest.approx(single, rel=0.12)
    assert m[0] > single * 0.999 - 4 * s[0]


def test_plain_and_roulette_local_estimates_agree(oracle):
    """Iwabuchi's Russian roulette must not change the expected upward radiance (8c-ix)."""
    d = fields.step_cloud(1.0)
    common = dict(surfaceAlbedo=0.1, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0])
    a = run_batches(make_integrator(oracle, d, useRussianRouletteForIntensity=False, **common), 15000, 8)
    b = run_batches(make_integrator(oracle, d, useRussianRouletteForIntensity=True, zetaMin=0.3, **common), 15000, 8, iseed=77)
    ma, sa = mean_se(a["meanIntensity"])
    mb, sb = mean_se(b["meanIntensity"])
    assert np.all(np.abs(ma - mb) < 4 * np.hypot(sa, sb))


def test_roulette_zeroes_downward_radiance_quirk_Q4(oracle):
    """With useRussianRouletteForIntensity the reference detects escape at the top only (MCRT:1555,1570,1583)."""
    I = make_integrator(oracle, fields.step_cloud(1.0), surfaceAlbedo=0.0, intensityMus=[-0.5], intensityPhis=[0.0],
                        useRussianRouletteForIntensity=True, zetaMin=0.3)
    r = run_batches(I, 5000, 2)
    assert np.all(r["meanIntensity"] == 0)


def _independent_slab_mc(tau, ssa, g, mu0, albedo, n, seed):
    """An independent plane-parallel Monte Carlo (numpy, float64, analytic Henyey-Greenstein sampling, textbook
    rotation of the direction) with the reference's tally definitions: fluxUp = weight leaving the top, fluxDown =
    weight arriving at the surface (every arrival), fluxAbsorbed = weight lost at collisions; no roulette."""
    rng = np.random.default_rng(seed)
    z = np.full(n, tau)                      # optical height above the surface
    mu = np.full(n, -mu0)
    phi = np.zeros(n)
    w = np.ones(n)
    alive = np.ones(n, bool)
    up = down = absorbed = 0.0
    nadir = 0.0  # weight leaving the top within the cone mu > 0.95
    while alive.any():
        i = np.flatnonzero(alive)
        s = -np.log(1.0 - rng.random(i.size))
        znew = z[i] + mu[i] * s
        out_top = znew >= tau
        hit = znew <= 0.0
        up += w[i[out_top]].sum()
        nadir += w[i[out_top]][mu[i[out_top]] > 0.95].sum()
        alive[i[out_top]] = False
        j = i[hit]                           # Lambertian reflection
        down += w[j].sum()
        w[j] *= albedo
        z[j] = 0.0
        mu[j] = np.sqrt(rng.random(j.size))
        phi[j] = 2 * np.pi * rng.random(j.size)
        alive[j[w[j] <= 0]] = False
        k = i[~out_top & ~hit]               # collisions
        z[k] = znew[~out_top & ~hit]
        absorbed += (w[k] * (1 - ssa)).sum()
        w[k] *= ssa
        xi = rng.random(k.size)
        cost = (1 + g * g - ((1 - g * g) / (1 - g + 2 * g * xi)) ** 2) / (2 * g)
        psi = 2 * np.pi * rng.random(k.size)
        sint, mu_k = np.sqrt(np.maximum(0, 1 - cost**2)), mu[k]
        sin0 = np.sqrt(np.maximum(1e-300, 1 - mu_k**2))
        mu[k] = np.clip(mu_k * cost + sin0 * sint * np.cos(psi), -1, 1)  # (azimuth is irrelevant in a slab)
        alive[k[w[k] < 1e-12]] = False
    return up / n, down / n, absorbed / n, nadir / n


@pytest.mark.parametrize("tau,ssa,albedo", [(1.0, 1.0, 0.0), (4.0, 0.9, 0.3)])
def test_multiple_scattering_slab_against_an_independent_monte_carlo(oracle, tau, ssa, albedo):
    """The whole chain of the oracle (inverse-table sampling, Marchuk rotation, ray tracing, surface reflection, implicit
    capture, roulette) against an independent textbook Monte Carlo of the same slab."""
    from tests.cases import make_integrator, run_batches, mean_se
    d = fields.plane_parallel(opticalDepth=tau, SSA=ssa, nLayers=3)
    I = make_integrator(oracle, d, surfaceAlbedo=albedo)
    r = run_batches(I, 20000, 12, source=dict(solarMu=0.5, solarAzimuth=0.0), want=["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed"])
    ind = np.array([_independent_slab_mc(tau, ssa, 0.85, 0.5, albedo, 40000, 100 + b) for b in range(6)])
    for col, key in enumerate(("meanFluxUp", "meanFluxDown", "meanFluxAbsorbed")):
        m, s = mean_se(r[key])
        mi, si = ind[:, col].mean(), ind[:, col].std(ddof=1) / np.sqrt(ind.shape[0])
        assert abs(m - mi) <= 3.5 * np.hypot(s, si) + 1e-6, (key, m, mi, s, si)


def test_nadir_radiance_normalisation_against_photon_binning(oracle):
    """The local estimate's normalisation (P/(4 pi |mu|), MCRT:1509; radiance per unit incident flux on the horizontal) against
    plain binning of the photons that leave an independent slab Monte Carlo within 18 degrees of the zenith:
    <I> ~ weight / (N * <mu> * solid angle), compared with the azimuthal mean of the oracle's radiances in the middle of that
    cone (the radiance varies by some 20 % across the cone, so the match is to a few percent)."""
    from tests.cases import make_integrator, run_batches, mean_se
    d = fields.plane_parallel(opticalDepth=2.0, SSA=1.0, nLayers=2)
    phis = [0.0, 60.0, 120.0, 180.0, 240.0, 300.0]
    I = make_integrator(oracle, d, surfaceAlbedo=0.0, intensityMus=[0.975] * 6, intensityPhis=phis, useRussianRouletteForIntensity=False)
    r = run_batches(I, 20000, 8, source=dict(solarMu=0.5, solarAzimuth=0.0), want=["meanIntensity"])
    per_batch = r["meanIntensity"].mean(axis=1)  # azimuthal mean, per batch
    m, s = per_batch.mean(), per_batch.std(ddof=1) / np.sqrt(per_batch.size)
    ind = np.array([_independent_slab_mc(2.0, 1.0, 0.85, 0.5, 0.0, 150000, 200 + b)[3] for b in range(4)])
    binned = ind / (0.975 * 2 * np.pi * 0.05)  # <mu> = 0.975 over mu in (0.95, 1]; solid angle 2 pi * 0.05
    bm, bs = binned.mean(), binned.std(ddof=1) / 2.0
    assert abs(m - bm) <= 0.03 * bm + 3 * np.hypot(s, bs), (m, s, bm, bs)


def test_translation_invariance_of_the_periodic_domain(oracle):
    from tests.cases import assert_translation_invariance
    assert_translation_invariance(oracle, fields.step_cloud(0.99), 5, 0,
                                  dict(surfaceAlbedo=0.2, intensityMus=[1.0, 0.5], intensityPhis=[0.0, 180.0], useRussianRouletteForIntensity=False),
                                  6000, 12, source=dict(solarMu=0.5, solarAzimuth=0.0))
