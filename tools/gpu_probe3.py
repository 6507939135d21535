#!/usr/bin/env python
"""All bench workloads once (development aid): throughput + sanity of results."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from bench import make_workload
from i3rc_monte_carlo_model_b200._lib import backend
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import getCounters, new_Integrator, reportResults, specifyParameters

be = backend()
for name, nph in (("planeparallel", 4_000_000), ("step", 2_000_000), ("radar", 1_000_000), ("landsat", 4_000_000),
                  ("les-small", 1_000_000), ("les", 1_000_000)):
    wl = make_workload(name)
    t0 = time.time()
    d = wl["domain"]()
    t1 = time.time()
    I = new_Integrator(d, backend=be)
    specifyParameters(I, **wl["params"])
    src = new_PhotonStream(numberOfPhotons=nph, **wl["source"]).as_c()
    be.stats_reset(I.handle, 0)
    assert be.run_batches(I.handle, C.byref(src), 10, 0, 100, 1) == 0, I._msg()
    t2 = time.time()
    be.reset_timing(I.handle)
    be.stats_reset(I.handle, 0)
    assert be.run_batches(I.handle, C.byref(src), 10, 0, 1, 2) == 0, I._msg()
    ms, nl, no = C.c_double(), C.c_int64(), C.c_int64()
    be.get_timing(I.handle, C.byref(ms), C.byref(nl), C.byref(no))
    c = getCounters(I)
    r = reportResults(I, "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity")
    n = nph * 2
    cross = c["crossings_photon"] + c["crossings_intensity"]
    print(f"{name:14s} build {t1-t0:.1f}s setup {t2-t1:.1f}s  {n/ms.value*1e3:.4g} ph/s {cross/ms.value*1e3:.4g} cross/s cross/ph={cross/n:.0f} "
          f"coll/ph={c['collisions']/n:.1f} contrib/ph={c['contributions']/n:.1f} bad={c['bad']}  closure="
          f"{r['meanFluxUp'] + r['meanFluxAbsorbed'] + (1 - wl['params']['surfaceAlbedo']) * r['meanFluxDown']:.4f} I={np.round(r['meanIntensity'][:4], 4)}", flush=True)
    del I
