#!/usr/bin/env python
"""Tuning probe for the transport kernel on the Landsat radiance workload (development aid)."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from bench import make_workload
from i3rc_monte_carlo_model_b200._lib import backend
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import getCounters, new_Integrator, reportResults, specifyParameters


def run(wlname, nph, nb, tune):
    be = backend()
    wl = make_workload(wlname)
    I = new_Integrator(wl["domain"](), backend=be)
    specifyParameters(I, **wl["params"])
    for k, v in tune.items():
        assert be.set_tuning(I.handle, k.encode(), v) == 0, (k, v)
    src = new_PhotonStream(numberOfPhotons=nph, **wl["source"]).as_c()
    be.stats_reset(I.handle, 0)
    assert be.run_batches(I.handle, C.byref(src), 10, 0, 100, 1) == 0, I._msg()
    be.reset_timing(I.handle)
    be.stats_reset(I.handle, 0)
    assert be.run_batches(I.handle, C.byref(src), 10, 0, 1, nb) == 0, I._msg()
    ms, nl, no = C.c_double(), C.c_int64(), C.c_int64()
    be.get_timing(I.handle, C.byref(ms), C.byref(nl), C.byref(no))
    c = getCounters(I)
    r = reportResults(I, "meanFluxUp", "meanIntensity")
    n = nph * nb
    cross = c["crossings_photon"] + c["crossings_intensity"]
    print(f"{wlname:10s} {str(tune):70s} {n/ms.value*1e3:.4g} ph/s {cross/ms.value*1e3:.4g} cross/s  up={r['meanFluxUp']:.5f} I={np.round(r['meanIntensity'],5)}", flush=True)


if __name__ == "__main__":
    wls = sys.argv[1:] or ["landsat"]
    for wl in wls:
        nph = 2_000_000 if wl != "les" else 500_000
        for tune in ({"resident_blocks": 6}, {"resident_blocks": 6, "steps_per_event_phase": 6}, {"resident_blocks": 6, "steps_per_event_phase": 4},
                     {"resident_blocks": 5}, {"resident_blocks": 5, "steps_per_event_phase": 4}, {"resident_blocks": 4}, {"resident_blocks": 4, "steps_per_event_phase": 4},
                     {"event_threshold": 8}, {"event_threshold": 24}, {"event_threshold": 30}):
            run(wl, nph, 2, tune)
