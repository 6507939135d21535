#!/usr/bin/env python
"""Development aid: throughput of the transport kernel outside bench.py.

    python tools/gpu_probe.py workloads                       every bench workload once (sanity of results + rates)
    python tools/gpu_probe.py tune <workload> <photons> '{"resident_blocks": 5}' '{"min_running": 20}' ...
                                                              one workload, one line per tuning dict (i3rc_set_tuning)
    python tools/gpu_probe.py sizes ['{tuning}']              Landsat at 1, 4, 16, 64 M photons per launch (tail effect)
    python tools/gpu_probe.py ceilings                        measured roofline ceilings: random 4-byte gathers per second
                                                              over arrays of 1 MB .. 1 GiB, and the issue rate
"""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from bench import make_workload
from i3rc_monte_carlo_model_b200._lib import backend
from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import getCounters, new_Integrator, reportResults, specifyParameters


def run(wlname, nph, nb, tune, verbose=False):
    be = backend()
    wl = make_workload(wlname)
    t0 = time.time()
    I = new_Integrator(wl["domain"](), backend=be)
    specifyParameters(I, **wl["params"])
    for k, v in tune.items():
        assert be.set_tuning(I.handle, k.encode(), v) == 0, (k, v)
    src = new_PhotonStream(numberOfPhotons=nph, **wl["source"]).as_c()
    be.stats_reset(I.handle, 0)
    assert be.run_batches(I.handle, C.byref(src), 10, 0, 100, 1) == 0, I._msg()  # warm-up (tables)
    t1 = time.time()
    be.reset_timing(I.handle)
    be.stats_reset(I.handle, 0)
    assert be.run_batches(I.handle, C.byref(src), 10, 0, 1, nb) == 0, I._msg()
    ms, nl, no = C.c_double(), C.c_int64(), C.c_int64()
    be.get_timing(I.handle, C.byref(ms), C.byref(nl), C.byref(no))
    c = getCounters(I)
    r = reportResults(I, "meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "meanIntensity")
    n = nph * nb
    cross = c["crossings_photon"] + c["crossings_intensity"]
    line = f"{wlname:13s} {str(tune):60s} {n/ms.value*1e3:.4g} ph/s {cross/ms.value*1e3:.4g} cross/s"
    if verbose:
        closure = r["meanFluxUp"] + r["meanFluxAbsorbed"] + (1 - wl["params"]["surfaceAlbedo"]) * r["meanFluxDown"]
        line += (f" setup {t1-t0:.1f}s cross/ph={cross/n:.0f} coll/ph={c['collisions']/n:.1f} contrib/ph={c['contributions']/n:.1f}"
                 f" bad={c['bad']} closure={closure:.4f}")
    print(line + f" up={r['meanFluxUp']:.5f} I={np.round(r['meanIntensity'][:4], 5)}", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "workloads"
    if mode == "workloads":
        for name, nph in (("planeparallel", 4_000_000), ("step", 2_000_000), ("radar", 1_000_000), ("landsat", 4_000_000),
                          ("les-small", 1_000_000), ("les", 1_000_000)):
            run(name, nph, 2, {}, verbose=True)
    elif mode == "ceilings":
        be = backend()
        g = C.c_double()
        for mb in (1, 8, 32, 64, 128, 256, 1024):
            assert be.measure_gather_rate(mb << 20, 5, C.byref(g)) == 0
            print(f"random 4-byte gathers over {mb:5d} MB: {g.value:.4g} loads/s = {g.value*4/1e9:.1f} GB/s useful, "
                  f"{g.value*32/1e9:.1f} GB/s of 32-byte sectors", flush=True)
        assert be.measure_issue_rate(5, C.byref(g)) == 0
        print(f"issue rate (independent FMAs): {g.value:.4g} warp instructions/s "
              f"(nominal 148 SMs x 4 schedulers x 1.965 GHz = {148*4*1.965e9:.4g})", flush=True)
    elif mode == "locality":
        # What would perfect gather locality buy?  A homogeneous domain of the Landsat case's shape and mean optical depth:
        # with all field strides zero every gather reads the same element (an L1 hit), and the results do not change.
        import bench
        from i3rc_monte_carlo_model_b200 import fields
        wl0 = bench.make_workload("landsat")
        bench_make = bench.make_workload
        def homogeneous(name):
            w = dict(wl0)
            w["domain"] = lambda: fields.plane_parallel(nX=128, nY=128, nLayers=119, domainSize=3840.0, physicalThickness=2380.0,
                                                        opticalDepth=10.0, nLegendreCoefficients=299)
            return w
        bench.make_workload = homogeneous
        globals()["make_workload"] = homogeneous
        for t in ({"vertical_shortcut": 0}, {"vertical_shortcut": 0, "debug_zero_strides": 1}, {}, {"debug_zero_strides": 1}):
            run("landsat-shaped homogeneous slab", 8_000_000, 1, t, verbose=True)
        bench.make_workload = bench_make
    elif mode == "tune":
        for t in sys.argv[4:] or ["{}"]:
            run(sys.argv[2], int(sys.argv[3]), 1, eval(t))
    elif mode == "sizes":
        for t in sys.argv[2:] or ["{}"]:
            for nph in (1_000_000, 4_000_000, 16_000_000, 64_000_000):
                run("landsat", nph, 1, eval(t))
