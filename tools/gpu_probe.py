#!/usr/bin/env python
"""Quick GPU probe: time the transport kernel on a few configurations and print counters (development aid)."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from i3rc_monte_carlo_model_b200 import _abi, fields
from i3rc_monte_carlo_model_b200._lib import backend
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (getCounters, new_Integrator, reportResults,
                                                                    specifyParameters)


def timing(I):
    ms, nl, no = C.c_double(), C.c_int64(), C.c_int64()
    I.backend.get_timing(I.handle, C.byref(ms), C.byref(nl), C.byref(no))
    return ms.value, nl.value, no.value


def run(name, d, nph, nb, tune=None, **kw):
    be = backend()
    I = new_Integrator(d, backend=be)
    specifyParameters(I, **kw)
    for k, v in (tune or {}).items():
        assert be.set_tuning(I.handle, k.encode(), v) == 0
    ph = new_PhotonStream(0.5, 0.0, numberOfPhotons=nph)
    src = ph.as_c()
    be.stats_reset(I.handle, 0)
    rc = be.run_batches(I.handle, C.byref(src), 10, 0, 1, 1)  # warm-up (tabulates)
    assert rc == 0, I._msg()
    be.reset_timing(I.handle)
    be.stats_reset(I.handle, 0)
    t0 = time.time()
    rc = be.run_batches(I.handle, C.byref(src), 10, 0, 1, nb)
    assert rc == 0, I._msg()
    wall = time.time() - t0
    ms, nl, no = timing(I)
    c = getCounters(I)
    r = reportResults(I, "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", *(["meanIntensity"] if "intensityMus" in kw else []))
    n = nph * nb
    cross = c["crossings_photon"] + c["crossings_intensity"]
    print(f"{name:28s} tune={tune} photons={n:.3g} kernel={ms:.1f}ms wall={wall*1e3:.1f}ms  {n/ms*1e3:.4g} ph/s  "
          f"{cross/ms*1e3:.4g} crossings/s  cross/ph={cross/n:.1f} coll/ph={c['collisions']/n:.2f}", flush=True)
    print("    ", {k: (round(float(v), 5) if np.ndim(v) == 0 else np.round(v, 5)) for k, v in r.items()}, flush=True)
    return I


if __name__ == "__main__":
    dirs3 = dict(intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0])
    run("planeParallel flux", fields.plane_parallel(), 2_000_000, 4, surfaceAlbedo=0.0)
    run("planeParallel rad RR", fields.plane_parallel(), 1_000_000, 4, surfaceAlbedo=0.0, useRussianRouletteForIntensity=True, zetaMin=0.3, **dirs3)
    run("step flux", fields.step_cloud(0.99), 1_000_000, 4, surfaceAlbedo=0.0)
    run("step rad RR", fields.step_cloud(0.99), 500_000, 4, surfaceAlbedo=0.0, useRussianRouletteForIntensity=True, zetaMin=0.3, **dirs3)
    land = fields.landsat_cloud(1.0)
    for tune in (None, {"steps_per_event_phase": 4}, {"steps_per_event_phase": 16}, {"block_size": 256}, {"block_size": 64}):
        run("landsat flux", land, 2_000_000, 3, tune=tune, surfaceAlbedo=0.0)
    for tune in (None, {"steps_per_event_phase": 4}, {"steps_per_event_phase": 16}, {"block_size": 256}, {"block_size": 64}):
        run("landsat rad RR", land, 500_000, 3, tune=tune, surfaceAlbedo=0.0, useRussianRouletteForIntensity=True, zetaMin=0.3, **dirs3)
    run("landsat rad plain", land, 200_000, 3, surfaceAlbedo=0.0, useRussianRouletteForIntensity=False, **dirs3)
