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
}


def main(which=None):
    be = oracle_backend()
    for name, (make, params, source, nph, nb) in CASES.items():
        if which and name not in which:
            continue
        t0 = time.time()
        I = make_integrator(be, make(), **params)
        sums, cnt = run_batches(I, new_PhotonStream(numberOfPhotons=nph, **source), 10, nb, seedOrder=0, nThreads=0)
        out = {"photonsPerBatch": nph, "numBatches": nb}
        for k, s in sums.items():
            mean = s[0] / nb
            var = np.maximum(s[1] / nb - mean**2, 0.0) * nb / (nb - 1)
            out[k + "_mean"] = mean.astype(np.float32)
            out[k + "_se"] = np.sqrt(var / nb).astype(np.float32)
        for k, v in cnt.items():
            out["cnt_" + k] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: {nph * nb} photons in {time.time() - t0:.1f}s  meanFluxUp={out['meanFluxUp_mean']:.5f}")


if __name__ == "__main__":
    main(sys.argv[1:])
