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
    for b in range(first_batch, first_batch + nb):
        ph = new_PhotonStream(numberOfPhotons=nph, **source)
        computeRadiativeTransfer(I, new_RandomNumberSequence([iseed, b]), ph)
        r = reportResults(I, *want)
        for k in want:
            acc[k].append(np.array(r[k], dtype=np.float64))
    out = {k: np.stack(v) for k, v in acc.items()}
    out["counters"] = getCounters(I)
    return out


def mean_se(x):
    """(mean, standard error) over batches; passes an already summarised (mean, se) pair through."""
    if isinstance(x, tuple):
        return x
    nb = x.shape[0]
    return x.mean(0), x.std(0, ddof=1) / np.sqrt(nb)


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
        out[names.get(k, k)] = (mean, se)
    out["counters"] = cnt
    return out


def zscores(a, b, floor=0.0):
    """(mean_a - mean_b) / combined standard error, elementwise.  ``floor`` is an absolute error floor for
    entries whose batch variance is degenerate (e.g. exactly zero in both)."""
    ma, sa = mean_se(a)
    mb, sb = mean_se(b)
    s = np.sqrt(sa**2 + sb**2 + floor**2)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(s > 0, (ma - mb) / s, np.where(ma == mb, 0.0, np.inf))
    return z


def assert_statistical_parity(a, b, keys=None, sigma=SIGMA, per_column_sigma=4.5, label=""):
    """Domain means within `sigma`; per-column fields: no column beyond per_column_sigma (the expected maximum of
    thousands of unit normals) and a chi-square consistent with unit variance."""
    keys = keys or [k for k in a if k != "counters" and k in b]
    for k in keys:
        z = np.atleast_1d(zscores(a[k], b[k], floor=1e-7))
        if z.size <= 32:
            assert np.all(np.abs(z) <= sigma + (0.8 if z.size > 4 else 0.0)), f"{label}{k}: z = {z}"
        else:
            zz = z[np.isfinite(z)]
            assert np.abs(zz).max() <= per_column_sigma + 0.3 * np.log10(zz.size), f"{label}{k}: max |z| = {np.abs(zz).max()}"
            chi2 = np.mean(zz**2)
            assert chi2 < 1.0 + 6.0 * np.sqrt(2.0 / zz.size) + 0.35, f"{label}{k}: mean z^2 = {chi2}"


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
