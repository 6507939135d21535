#!/usr/bin/env python
"""Generates tests/golden/*.npz: batch means and standard errors of the CPU oracle for the larger benchmark
configurations, so that the `-m gpu` parity tests can compare at sizes the oracle does not finish in seconds.

The reference itself cannot run (no Fortran compiler; PARITY UNPINNED), so these are outputs of the oracle
restatement with the reference's MT19937 streams ((/ iseed, batch /), monteCarloDriver.f95:277).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from i3rc_monte_carlo_model_b200 import fields  # noqa: E402
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream  # noqa: E402
from oracle.binding import oracle_backend, run_batches  # noqa: E402
from tests.cases import make_integrator  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (domain factory, specifyParameters kwargs, source kwargs, photons per batch, batches)
    "landsat_rr": (lambda: fields.landsat_cloud(1.0), dict(
        surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0],
        useRussianRouletteForIntensity=True, zetaMin=0.3, minInverseTableSize=10001, minForwardTableSize=10001),
        dict(solarMu=0.5, solarAzimuth=0.0), 20000, 32),
    "landsat_absorbing_flux": (lambda: fields.landsat_cloud(0.99), dict(surfaceAlbedo=0.2),
                               dict(solarMu=0.8, solarAzimuth=45.0), 50000, 32),
    "radar_c1_plain": (lambda: fields.radar_cloud(0.99, "C1"), dict(
        surfaceAlbedo=0.1, intensityMus=[1.0, -0.6], intensityPhis=[0.0, 30.0], useRussianRouletteForIntensity=False),
        dict(solarMu=0.5, solarAzimuth=0.0), 4000, 32),
    "les_small_two_components": (lambda: fields.synthetic_les(nx=32, ny=32, nz=32, n_entries=5, seed=7), dict(
        surfaceAlbedo=0.05, intensityMus=[1.0, 0.4], intensityPhis=[0.0, 270.0], useRussianRouletteForIntensity=True,
        zetaMin=0.3), dict(solarMu=0.6, solarAzimuth=20.0), 10000, 32),
    # ---- round 2: high statistical power on the benchmark configuration, and the configurations round 1 left untested
    # the bench.py default workload at 6.4e7 oracle photons: sigma(meanFluxUp) ~ 1e-4 instead of 1e-3
    "landsat_rr_hi": (lambda: fields.landsat_cloud(1.0), dict(
        surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0],
        useRussianRouletteForIntensity=True, zetaMin=0.3, minInverseTableSize=10001, minForwardTableSize=10001),
        dict(solarMu=0.5, solarAzimuth=0.0), 1000000, 64),
    # ... and 64 more batches (numbers 65-128, independent streams) of the same: the test combines both fixtures
    "landsat_rr_hi_b": (lambda: fields.landsat_cloud(1.0), dict(
        surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0],
        useRussianRouletteForIntensity=True, zetaMin=0.3, minInverseTableSize=10001, minForwardTableSize=10001),
        dict(solarMu=0.5, solarAzimuth=0.0), 1000000, 64),
    # BASELINE config 4 (radar cloud, Deirmendjian C1, absorbing) with Russian roulette for intensity
    "radar_c1_rr": (lambda: fields.radar_cloud(0.99, "C1"), dict(
        surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0],
        useRussianRouletteForIntensity=True, zetaMin=0.3, minInverseTableSize=10001, minForwardTableSize=10001),
        dict(solarMu=0.5, solarAzimuth=0.0), 500000, 64),
    # BASELINE config 2 (step cloud) with the sun overhead
    "step_mu1_rr": (lambda: fields.step_cloud(0.99), dict(
        surfaceAlbedo=0.0, intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0],
        useRussianRouletteForIntensity=True, zetaMin=0.3), dict(solarMu=1.0, solarAzimuth=0.0), 1000000, 64),
    # BASELINE config 5 scaled to 256x256x200 (52 MB extinction field: large enough for new_Integrator to choose the
    # layer-compacted field and the SPLIT kernel by itself), 2 components, 27-entry table, 16 directions
    "les_mid_split": (lambda: fields.synthetic_les(nx=256, ny=256, nz=200), dict(
        surfaceAlbedo=0.05, intensityMus=[m for m in (1.0, 0.8, 0.6, 0.4) for _ in range(4)],
        intensityPhis=[p for _ in range(4) for p in (0.0, 90.0, 180.0, 270.0)], useRussianRouletteForIntensity=True,
        zetaMin=0.3, minInverseTableSize=10001, minForwardTableSize=10001),
        dict(solarMu=0.5, solarAzimuth=30.0), 50000, 64),
}
BATCH_BEGIN = {"landsat_rr_hi_b": 65}  # first batch number (default 1): seeds are (/ iseed, batch /)
COARSEN = {"les_mid_split": 8}  # per-column fields stored as means over 8x8 blocks of columns (fixture size)


def coarsen_mean_se(mean, se, f):
    """[..., ny, nx] per-column mean and standard error -> blocks of f x f columns.  The block standard error treats the
    columns of a block as independent (their batch covariance is not kept by the driver's moments)."""
    ny, nx = mean.shape[-2:]
    lead = mean.shape[:-2]
    m = mean.reshape(*lead, ny // f, f, nx // f, f).mean(axis=(-3, -1))
    v = (se**2).reshape(*lead, ny // f, f, nx // f, f).sum(axis=(-3, -1)) / f**4
    return m, np.sqrt(v)


def main(which=None):
    be = oracle_backend()
    for name, (make, params, source, nph, nb) in CASES.items():
        if which and name not in which:
            continue
        t0 = time.time()
        I = make_integrator(be, make(), **params)
        sums, cnt = run_batches(I, new_PhotonStream(numberOfPhotons=nph, **source), 10, nb, seedOrder=0, nThreads=0,
                                batchBegin=BATCH_BEGIN.get(name, 1))
        out = {"photonsPerBatch": nph, "numBatches": nb}
        for k, s in sums.items():
            mean = s[0] / nb
            var = np.maximum(s[1] / nb - mean**2, 0.0) * nb / (nb - 1)
            se = np.sqrt(var / nb)
            if name in COARSEN and mean.ndim >= 2:
                mean, se = coarsen_mean_se(mean, se, COARSEN[name])
                out["coarsen"] = COARSEN[name]
            out[k + "_mean"] = mean.astype(np.float32)
            out[k + "_se"] = se.astype(np.float32)
        for k, v in cnt.items():
            out["cnt_" + k] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: {nph * nb} photons in {time.time() - t0:.1f}s  meanFluxUp={out['meanFluxUp_mean']:.5f}")


if __name__ == "__main__":
    main(sys.argv[1:])
