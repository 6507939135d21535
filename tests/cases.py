"""Shared helpers of the parity tests: the same call sequence (the reference's driver loop,
Example-Drivers/monteCarloDriver.f95:274-326) is run on a candidate backend and on the oracle, and compared
within a stated multiple of the combined Monte Carlo batch standard error."""
from __future__ import annotations

import numpy as np

from i3rc_monte_carlo_model_b200 import fields
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, getCounters, new_Integrator,
                                                                    reportResults, specifyParameters)
from i3rc_monte_carlo_model_b200.RandomNumbers import new_RandomNumberSequence

SIGMA = 3.0  # the tolerance BASELINE.json states: 3 sigma of the combined batch standard error


def make_integrator(backend, domain, **params):
    I = new_Integrator(domain, backend=backend)
    assert I.handle, "new_Integrator failed"
    specifyParameters(I, **params)
    return I


def run_batches(I, nph, nb, source=None, iseed=10, first_batch=1, want=None):
    """Returns dict name -> array [nb, ...] of per-batch results, plus counters of the last batch."""
    source = source or dict(solarMu=0.5, solarAzimuth=0.0)
    want = want or (["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxUp", "fluxDown", "fluxAbsorbed", "absorbedProfile"]
                    + (["meanIntensity", "intensity"] if I.nDir else []))
    acc = {k: [] for k in want}
    cnts = []
    for b in range(first_batch, first_batch + nb):
        ph = new_PhotonStream(numberOfPhotons=nph, **source)
        computeRadiativeTransfer(I, new_RandomNumberSequence([iseed, b]), ph)
        r = reportResults(I, *want)
        for k in want:
            acc[k].append(np.array(r[k], dtype=np.float64))
        cnts.append(getCounters(I))
    out = {k: np.stack(v) for k, v in acc.items()}
    out["counters"] = cnts[-1]
    out["counters_batches"] = {k: np.array([c[k] for c in cnts], np.float64) for k in cnts[-1]}
    return out


def assert_counter_parity(got, ref_counters, names, label="", sigma=SIGMA, floor=1e-3):
    """Events per photon (collisions, cell crossings, ...) of the candidate, batch by batch, against the reference
    side's totals: a z test with the candidate's batch-to-batch scatter standing for both sides (the oracle's batch
    driver keeps totals only), held to the family-wise 3-sigma bound.  ``floor``: relative error floor for counters whose
    definition leaves room for round-off-level differences (a ray that ends on a cell face)."""
    cb = got["counters_batches"]
    nb = cb["photons"].size
    nb_ref = max(ref_counters["photons"] / max(cb["photons"].mean(), 1.0), 1.0)
    zs = []
    for c in names:
        events = cb[c]
        if c == "crossings_photon" and "cells_skipped" in cb:  # cells crossed without a look count as crossed
            events = events + cb["cells_skipped"]
        rate = events / np.maximum(cb["photons"], 1.0)
        a, sd = rate.mean(), rate.std(ddof=1)
        b = ref_counters[c] / max(ref_counters["photons"], 1)
        zs.append((a - b) / np.sqrt(sd**2 * (1.0 / nb + 1.0 / nb_ref) + (floor * b) ** 2 + 1e-30))
    bound = familywise_bound(len(names), nb - 1, sigma)
    assert np.all(np.abs(zs) <= bound), f"{label}events per photon {list(names)}: z = {np.round(zs, 2)} (bound {bound:.2f})"


def mean_se_n(x):
    """(mean, standard error, number of batches) over batches; passes an already summarised (mean, se[, nb]) through."""
    if isinstance(x, tuple):
        return (x[0], x[1], x[2] if len(x) > 2 else 32)
    nb = x.shape[0]
    return x.mean(0), x.std(0, ddof=1) / np.sqrt(nb), nb


def mean_se(x):
    return mean_se_n(x)[:2]


def familywise_bound(m, dof=np.inf, sigma=SIGMA):
    """|z| bound for a FAMILY of m comparisons such that the family as a whole raises a false alarm as often as ONE
    two-sided `sigma` test of a normal variable does (Bonferroni).  The statistic is a difference of batch means over
    a standard error ESTIMATED from a few dozen batches, i.e. Student-t distributed with `dof` degrees of freedom
    (Welch-Satterthwaite), so the quantile is taken from that distribution.  m = 1, dof = inf gives exactly `sigma`."""
    from scipy import stats
    alpha = 2.0 * stats.norm.sf(sigma)
    q = alpha / (2.0 * max(int(m), 1))
    return float(stats.norm.isf(q) if not np.isfinite(dof) else stats.t.isf(q, dof))


def oracle_summary(I, nph, nb, source=None, iseed=10):
    """The oracle side of a parity test through its OpenMP batch driver (threads stand in for MPI ranks):
    name -> (mean, standard error), arrays indexed like reportResults ([x, y(, direction)])."""
    from oracle.binding import run_batches as oracle_run_batches
    source = source or dict(solarMu=0.5, solarAzimuth=0.0)
    sums, cnt = oracle_run_batches(I, new_PhotonStream(numberOfPhotons=nph, **source), iseed, nb, seedOrder=0, nThreads=0)
    out = {}
    names = {"radiance": "intensity", "meanRadiance": "meanIntensity"}
    for k, s in sums.items():
        mean = s[0] / nb
        se = np.sqrt(np.maximum(s[1] / nb - mean**2, 0.0) / (nb - 1))
        if mean.ndim >= 2:
            mean, se = mean.T, se.T
        out[names.get(k, k)] = (mean, se, nb)
    out["counters"] = cnt
    return out


def zscores(a, b, floor=0.0, with_dof=False):
    """(mean_a - mean_b) / combined standard error, elementwise (Welch's t statistic).  ``floor`` is an absolute error
    floor for entries whose batch variance is degenerate (e.g. exactly zero in both).  with_dof: also the
    Welch-Satterthwaite degrees of freedom of the statistic (the smallest over the elements)."""
    ma, sa, na = mean_se_n(a)
    mb, sb, nb = mean_se_n(b)
    s = np.sqrt(sa**2 + sb**2 + floor**2)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(s > 0, (ma - mb) / s, np.where(ma == mb, 0.0, np.inf))
        if with_dof:
            dof = np.where(s > 0, (sa**2 + sb**2) ** 2 / (sa**4 / (na - 1) + sb**4 / (nb - 1) + 1e-300), np.inf)
            return z, float(max(np.min(dof), min(na, nb) - 1))
    return z


def assert_statistical_parity(a, b, keys=None, sigma=SIGMA, label=""):
    """BASELINE.json's criterion: 3 sigma of the combined batch standard error.  A quantity that is a vector of m values
    (three radiances, 16384 columns) is a family of m comparisons and is held to the FAMILY-WISE 3-sigma bound
    (familywise_bound): the vector as a whole fails as often as a single 3-sigma test would, whatever its length.
    Per-column fields must also have a mean square of z consistent with its expectation dof/(dof-2) under parity:
    within 6 standard deviations of the mean of m such squares, plus 0.15 for the correlation between neighbouring
    columns within a batch, which the per-column standard errors do not see."""
    keys = keys or [k for k in a if not k.startswith("counters") and k in b]
    for k in keys:
        z, dof = zscores(a[k], b[k], floor=1e-7, with_dof=True)
        zz = np.atleast_1d(z)
        zz = zz[np.isfinite(zz)] if zz.size > 32 else zz
        bound = familywise_bound(zz.size, dof, sigma)
        assert np.abs(zz).max() <= bound, f"{label}{k}: max |z| = {np.abs(zz).max():.2f} > {bound:.2f} (family of {zz.size}, dof {dof:.0f})\n{zz if zz.size <= 32 else ''}"
        if zz.size > 32:
            chi2, expect = np.mean(zz**2), dof / (dof - 2.0)
            spread = expect * np.sqrt(2.0 * (dof - 1.0) / (dof - 4.0) / zz.size)
            assert chi2 < expect + 6.0 * spread + 0.15, f"{label}{k}: mean z^2 = {chi2:.3f} (expected {expect:.3f})"


CONFIGS = {
    "planeParallel": lambda: fields.plane_parallel(),
    "stepCloud": lambda: fields.step_cloud(1.0),
    "stepCloudAbsorbing": lambda: fields.step_cloud(0.99),
}


def rolled_domain(d, kx, ky):
    """The same periodic domain shifted by kx columns in x and ky in y (every component rolled)."""
    from i3rc_monte_carlo_model_b200.opticalProperties import addOpticalComponent, new_Domain
    out = new_Domain(d.xPosition, d.yPosition, d.zPosition)
    for c in d.components:
        if c.horizontallyUniform:
            e, s, p = c.extinction, c.singleScatteringAlbedo, c.phaseFunctionIndex
        else:
            e, s, p = (np.roll(a, (kx, ky), axis=(0, 1)) for a in (c.extinction, c.singleScatteringAlbedo, c.phaseFunctionIndex))
        addOpticalComponent(out, c.name, e, s, p, c.table, zLevelBase=c.zLevelBase)
    return out


def assert_translation_invariance(backend, d, kx, ky, params, nph, nb, source=None):
    """Periodic boundaries: shifting the domain shifts the per-column results and nothing else (a property that needs no
    reference values; it exercises cell indexing, the periodic wrap and the exit-column bookkeeping)."""
    a = run_batches(make_integrator(backend, d, **params), nph, nb, source=source)
    b = run_batches(make_integrator(backend, rolled_domain(d, kx, ky), **params), nph, nb, source=source, iseed=77)
    for k in ("fluxUp", "fluxDown", "fluxAbsorbed", "intensity"):
        if k not in a:
            continue
        rolled = np.roll(a[k], (kx, ky), axis=(1, 2))
        assert_statistical_parity({k: rolled}, {k: b[k]}, keys=[k], label=f"shift ({kx},{ky}): ")
