"""The caller of the hot path: the batch loop, batch partitioning and moment statistics of
Example-Drivers/monteCarloDriver.f95 (:264-378), and its namelist reader (:90-103, 143-150).

Two ways through the loop:
  * ``run_batches_host``  -- the reference's own call sequence per batch (new_PhotonStream,
    computeRadiativeTransfer, reportResults into HOST arrays, moments on the host).  Works with any backend.
  * ``run_batches_device`` -- the same loop through ``i3rc_run_batches``: photons are generated, traced,
    normalised and folded into the sum(x) / sum(x*x) buffers on the device; nothing crosses PCIe per batch.
Ranks (one process per GPU) take contiguous blocks of batches exactly like monteCarloDriver.f95:264-274 and the
moment buffers are summed with ONE all-reduce (replacing the nine MPI_REDUCE calls of :333-348).
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np

from . import _abi
from .monteCarloIllumination import new_PhotonStream
from .monteCarloRadiativeTransfer import computeRadiativeTransfer, reportResults
from .RandomNumbers import new_RandomNumberSequence


def partition_batches(numBatches, numProcs=1, thisProc=0):
    """monteCarloDriver.f95:264-274: at least two batches, rounded up to a multiple of the number of processes;
    process p takes batches p*bpp+1 .. (p+1)*bpp.  Returns (numBatches, range of this process)."""
    numBatches = max(int(numBatches), 2)
    bpp = numBatches // numProcs
    if numBatches % numProcs != 0:
        bpp += 1
        numBatches = bpp * numProcs
    return numBatches, range(thisProc * bpp + 1, thisProc * bpp + bpp + 1)


class BatchStatistics:
    """First and second moments of every output over batches (monteCarloDriver.f95:300-321), in float64."""

    def __init__(self):
        self.sums: dict[str, np.ndarray] = {}
        self.nLocal = 0

    def add(self, results: dict):
        for k, v in results.items():
            v = np.asarray(v, dtype=np.float64)
            if k not in self.sums:
                self.sums[k] = np.zeros((2,) + v.shape)
            self.sums[k][0] += v
            self.sums[k][1] += v * v
        self.nLocal += 1

    def pack(self):
        keys = sorted(self.sums)
        return keys, np.concatenate([self.sums[k].ravel() for k in keys]) if keys else np.zeros(0)

    def unpack(self, keys, flat):
        o = 0
        for k in keys:
            n = self.sums[k].size
            self.sums[k] = flat[o:o + n].reshape(self.sums[k].shape).copy()
            o += n

    def allreduce(self, dist=None):
        """sumAcrossProcesses (Code/multipleProcesses_mpi.f95:57-131) as one all-reduce of one packed buffer."""
        if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
            return
        import torch
        keys, flat = self.pack()
        t = torch.from_numpy(flat)
        dist.all_reduce(t)
        self.unpack(keys, t.numpy())

    def finish(self, solarFlux, numBatches):
        """mean = solarFlux*sum/nB ; stderr = sqrt(max(0, solarFlux*sum2/nB - mean^2)/(nB-1))  (:358-378)"""
        out = {}
        for k, s in self.sums.items():
            mean = solarFlux * s[0] / numBatches
            second = solarFlux * s[1] / numBatches
            out[k] = (mean, np.sqrt(np.maximum(0.0, second - mean**2) / (numBatches - 1)))
        return out


def run_batches_host(integ, source, numPhotonsPerBatch, batches, iseed=10, seedOrder=0, want=None, stats=None):
    """monteCarloDriver.f95:274-326, one call per reference call; ``source`` holds new_PhotonStream's arguments."""
    stats = stats or BatchStatistics()
    want = want or (["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxUp", "fluxDown", "fluxAbsorbed",
                     "absorbedProfile"] + (["intensity", "meanIntensity"] if integ.nDir else []))
    for batch in batches:
        seed = [iseed, batch] if seedOrder == 0 else [batch, iseed]
        randoms = new_RandomNumberSequence(seed)
        photons = new_PhotonStream(numberOfPhotons=numPhotonsPerBatch, randomNumbers=randoms, **source)
        computeRadiativeTransfer(integ, randoms, photons)
        stats.add(reportResults(integ, *want))
    return stats


class _DeviceArray:
    """__cuda_array_interface__ view of the library's packed moment buffer (for torch.distributed)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def run_batches_device(integ, source, numPhotonsPerBatch, batches, iseed=10, seedOrder=0, with_volume=False, reset=True):
    """The same loop without host round trips (i3rc_run_batches).  Returns the counters of the run."""
    be = integ.backend
    if be.run_batches is None:
        raise RuntimeError("this backend has no device batch loop")
    src = new_PhotonStream(numberOfPhotons=numPhotonsPerBatch, **source).as_c()
    if reset:
        rc = be.stats_reset(integ.handle, int(with_volume))
        if rc == _abi.FAILURE:
            raise RuntimeError(integ._msg())
    batches = list(batches)
    assert batches == list(range(batches[0], batches[0] + len(batches)))
    rc = be.run_batches(integ.handle, C.byref(src), int(iseed), int(seedOrder), batches[0], len(batches))
    if rc == _abi.FAILURE:
        raise RuntimeError(integ._msg())
    c = _abi.Counters()
    be.get_counters(integ.handle, C.byref(c))
    return c.as_dict()


def allreduce_device_stats(integ, dist=None, device=None):
    """ONE all-reduce (NCCL over NVLink) of the packed device moment buffer."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    import torch
    be = integ.backend
    ptr, n = C.c_void_p(), C.c_int64()
    if be.stats_device_buffer(integ.handle, C.byref(ptr), C.byref(n)) == _abi.FAILURE:
        raise RuntimeError("no batch statistics to reduce")
    be.synchronize(integ.handle)
    t = torch.as_tensor(_DeviceArray(ptr.value, n.value), device=device or torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t)
    torch.cuda.synchronize()


def device_stats_report(integ, solarFlux, numBatches, with_volume=False):
    """monteCarloDriver.f95:358-378 on the device; returns name -> (mean, stderr) arrays indexed [x, y(, ..)]."""
    be = integ.backend
    nx, ny, nz, nd = integ.nx, integ.ny, integ.nz, integ.nDir
    out = _abi.StatsOut()
    bufs = {
        "meanFluxUp": np.zeros(2), "meanFluxDown": np.zeros(2), "meanFluxAbsorbed": np.zeros(2),
        "fluxUp": np.zeros((2, ny, nx)), "fluxDown": np.zeros((2, ny, nx)), "fluxAbsorbed": np.zeros((2, ny, nx)),
        "absorbedProfile": np.zeros((2, nz)),
    }
    if nd:
        bufs["radiance"] = np.zeros((2, nd, ny, nx))
        bufs["meanRadiance"] = np.zeros((2, nd))
    if with_volume:
        bufs["absorbedVolume"] = np.zeros((2, nz, ny, nx))
    for k, a in bufs.items():
        setattr(out, k, _abi.dptr(a))
    rc = be.stats_report(integ.handle, float(solarFlux), int(numBatches), C.byref(out))
    if rc == _abi.FAILURE:
        raise RuntimeError(integ._msg())
    res = {}
    for k, a in bufs.items():
        mean, err = a[0], a[1]
        if mean.ndim >= 2:  # [.., y, x] -> Fortran-like [x, y, ..]
            mean, err = mean.T, err.T
        res[k] = (mean, err)
    return res


# ---- spectral loop -------------------------------------------------------------------------------
def run_spectral_bands(integ, source, numPhotonsPerBatch, numBatches, gasComponent, bands, iseed=10, solarFlux=1.0, dist=None):
    """A k-distribution / band loop over ONE integrator: for every term ``(weight, extinctionProfile[nZ])`` the gas
    component's extinction is swapped on the device (``setComponentProfile``, no field rebuild), the usual batches are
    traced, and the per-term means are combined: mean = sum_k w_k mean_k, stderr = sqrt(sum_k w_k^2 stderr_k^2) (terms are
    independent).  The reference has no such loop (Code/kDistribution.f95 is an unfinished stub); its only mechanism for
    gas absorption is an extra ssa = 0 component (Tools/PhysicalPropertiesToDomain.f95:330-347), which is what is
    swapped here.  Returns name -> (mean, stderr)."""
    from .monteCarloRadiativeTransfer import setComponentProfile
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    nB, mine = partition_batches(numBatches, world, rank)
    total = None
    for k, (weight, profile) in enumerate(bands):
        setComponentProfile(integ, gasComponent, profile)
        # every term gets its own seeds: batches (k * nB + 1) ... of the same iseed
        run_batches_device(integ, source, numPhotonsPerBatch, [b + k * nB for b in mine], iseed=iseed)
        allreduce_device_stats(integ, dist)
        st = device_stats_report(integ, solarFlux, nB)
        if total is None:
            total = {name: (np.zeros_like(np.asarray(m, np.float64)), np.zeros_like(np.asarray(e, np.float64))) for name, (m, e) in st.items()}
        for name, (m, e) in st.items():
            total[name][0][...] += weight * np.asarray(m)
            total[name][1][...] += (weight * np.asarray(e)) ** 2
    return {name: (m, np.sqrt(v)) for name, (m, v) in total.items()}


# ---- namelists (monteCarloDriver.f95:90-103) ------------------------------------------------------------------
_DEFAULTS = {
    "radiativetransfer": dict(solarFlux=1.0, solarMu=1.0, solarAzimuth=0.0, surfaceAlbedo=0.0, intensityMus=[], intensityPhis=[]),
    "montecarlo": dict(numPhotonsPerBatch=0, numBatches=100, iseed=10, nPhaseIntervals=10001),
    "algorithms": dict(useRayTracing=True, useRussianRoulette=True, useHybridPhaseFunsForIntenCalcs=False,
                       hybridPhaseFunWidth=7.0, numOrdersOrigPhaseFunIntenCalcs=0, useRussianRouletteForIntensity=True,
                       zetaMin=0.3, limitIntensityContributions=False, maxIntensityContribution=77.0),
    "output": dict(reportVolumeAbsorption=False, reportAbsorptionProfile=False),
    "filenames": dict(domainFileName="", outputFluxFile="", outputRadFile="", outputAbsProfFile="",
                      outputAbsVolumeFile="", outputNetcdfFile=""),
}


def _parse_value(tok):
    t = tok.strip()
    if re.fullmatch(r"\.?(t|true)\.?", t, flags=re.I):
        return True
    if re.fullmatch(r"\.?(f|false)\.?", t, flags=re.I):
        return False
    if (t.startswith('"') and t.endswith('"')) or (t.startswith("'") and t.endswith("'")):
        return t[1:-1]
    try:
        return int(t)
    except ValueError:
        return float(t.replace("d", "e").replace("D", "e"))


def read_namelists(path):
    """A Fortran-namelist reader sufficient for the driver's five groups; returns {group: {name: value}} with the
    reference's defaults (monteCarloDriver.f95:61-88) filled in and names in the reference's spelling."""
    text = "\n".join(line.split("!", 1)[0] for line in open(path).read().splitlines())
    out = {g: dict(v) for g, v in _DEFAULTS.items()}
    for m in re.finditer(r"&(\w+)(.*?)(?:^|\s)/", text, flags=re.S | re.M):
        group = m.group(1).lower()
        if group not in out:
            continue
        canon = {k.lower(): k for k in out[group]}
        body = m.group(2)
        for am in re.finditer(r"(\w+)\s*=\s*(.*?)(?=(?:\w+\s*=)|\Z)", body, flags=re.S):
            name, raw = am.group(1).lower(), am.group(2).strip().rstrip(",")
            if name not in canon:
                continue
            toks = [t for t in re.split(r"[,\s]+(?=(?:[^\"']*[\"'][^\"']*[\"'])*[^\"']*$)", raw) if t.strip()]
            vals = [_parse_value(t) for t in toks]
            out[group][canon[name]] = vals if isinstance(out[group][canon[name]], list) else vals[0]
    rt = out["radiativetransfer"]
    n = sum(1 for v in rt["intensityMus"] if abs(v) > 0)  # numRadDir = count(abs(intensityMus) > 0), :151
    rt["intensityMus"], rt["intensityPhis"] = rt["intensityMus"][:n], (rt["intensityPhis"] + [0.0] * n)[:n]
    return out


# ---- the whole program (Example-Drivers/monteCarloDriver.f95:137-419) ----------------------------------------------
def monteCarloDriver(namelistFileName, backend=None, dist=None, verbose=True):
    """What `monteCarloDriver namelist.nml` does, on the GPU: read the five namelists and the domain file, set the
    integrator up exactly like the reference (:169-216), trace one photon to build the tables (:240-253), run this
    rank's block of batches on the device (:264-326), sum the moments over ranks with ONE all-reduce (:333-348), and
    have rank 0 write the ASCII / netCDF result files in the reference's formats (:382-419).  One process per GPU;
    ``dist`` is an initialised torch.distributed (or None).  Returns name -> (mean, stderr) on every rank."""
    import time

    from . import fileIO
    from ._lib import backend as cuda_backend
    from .ErrorMessages import ErrorMessage, stateIsFailure, getCurrentMessage
    from .monteCarloRadiativeTransfer import new_Integrator, specifyParameters

    t0 = time.perf_counter()
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    nml = read_namelists(namelistFileName)
    cfg = {}
    for g in nml.values():
        cfg.update(g)
    numRadDir = len(cfg["intensityMus"])
    computeIntensity = numRadDir > 0 and bool(cfg["outputRadFile"] or cfg["outputNetcdfFile"])
    if not computeIntensity:
        cfg["outputRadFile"] = ""

    status = ErrorMessage()

    def check(what):  # printStatus: stop on failure (Code/userInterface_Unix.f95:21-54)
        if stateIsFailure(status):
            raise RuntimeError(f"{what}: {getCurrentMessage(status)}")

    thisDomain = fileIO.read_Domain(cfg["domainFileName"], status)
    check("read_Domain")
    x, y, z = thisDomain.xPosition.copy(), thisDomain.yPosition.copy(), thisDomain.zPosition.copy()
    I = new_Integrator(thisDomain, status=status, backend=backend or cuda_backend())
    check("new_Integrator")
    fileIO.finalize_Domain(thisDomain)
    specifyParameters(I, surfaceAlbedo=cfg["surfaceAlbedo"], minInverseTableSize=cfg["nPhaseIntervals"], status=status)
    check("specifyParameters")
    if computeIntensity:
        specifyParameters(I, minForwardTableSize=cfg["nPhaseIntervals"], intensityMus=cfg["intensityMus"],
                          intensityPhis=cfg["intensityPhis"], computeIntensity=True, status=status)
        check("specifyParameters")
    specifyParameters(I, useRayTracing=cfg["useRayTracing"], useRussianRoulette=cfg["useRussianRoulette"], status=status)
    check("specifyParameters")
    if computeIntensity:
        specifyParameters(I, useHybridPhaseFunsForIntenCalcs=cfg["useHybridPhaseFunsForIntenCalcs"],
                          hybridPhaseFunWidth=cfg["hybridPhaseFunWidth"],
                          numOrdersOrigPhaseFunIntenCalcs=cfg["numOrdersOrigPhaseFunIntenCalcs"],
                          useRussianRouletteForIntensity=cfg["useRussianRouletteForIntensity"], zetaMin=cfg["zetaMin"],
                          limitIntensityContributions=cfg["limitIntensityContributions"],
                          maxIntensityContribution=cfg["maxIntensityContribution"], status=status)
        check("specifyParameters")
    source = dict(solarMu=cfg["solarMu"], solarAzimuth=cfg["solarAzimuth"])
    # one photon with seed (/ iseed, 0 /): checks the set-up and builds the tables before the batches
    ph = new_PhotonStream(numberOfPhotons=1, **source)
    computeRadiativeTransfer(I, new_RandomNumberSequence([cfg["iseed"], 0]), ph, status=status)
    check("computeRadiativeTransfer")
    tSetup = time.perf_counter() - t0
    numBatches, mine = partition_batches(cfg["numBatches"], world, rank)
    cfg["numBatches"] = numBatches
    if verbose and rank == 0:
        print(f" Setup CPU time (secs, approx): {int(tSetup)}")
        print(f" Doing {len(mine)} batches on each of {world} processors.")
    withVolume = bool(cfg["reportVolumeAbsorption"] or cfg["outputAbsVolumeFile"])
    if I.backend.stats_reset is not None:  # the product: the whole batch loop stays on the device
        run_batches_device(I, source, cfg["numPhotonsPerBatch"], mine, iseed=cfg["iseed"], with_volume=withVolume)
        allreduce_device_stats(I, dist)
        stats = device_stats_report(I, cfg["solarFlux"], numBatches, with_volume=withVolume)
    else:  # a backend without a device batch loop (the test-suite's checker): the reference's own host loop
        want = (["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxUp", "fluxDown", "fluxAbsorbed", "absorbedProfile"]
                + (["volumeAbsorption"] if withVolume else []) + (["intensity", "meanIntensity"] if computeIntensity else []))
        hs = run_batches_host(I, source, cfg["numPhotonsPerBatch"], mine, iseed=cfg["iseed"], want=want)
        hs.allreduce(dist)
        names = {"intensity": "radiance", "meanIntensity": "meanRadiance", "volumeAbsorption": "absorbedVolume"}
        stats = {names.get(k, k): v for k, v in hs.finish(cfg["solarFlux"], numBatches).items()}
    tTotal = time.perf_counter() - t0
    if rank == 0:
        if verbose:
            print(f" Total CPU time (secs, approx): {int(tTotal)}")
        if any(cfg[k] for k in ("outputFluxFile", "outputAbsProfFile", "outputAbsVolumeFile", "outputRadFile")):
            fileIO.writeResults_ASCII(cfg, x, y, z, stats)
            if verbose:
                print(" Wrote ASCII results")
        if cfg["outputNetcdfFile"]:
            out = dict(stats)
            if not computeIntensity:
                out.pop("radiance", None)
            fileIO.writeResults_netcdf(cfg, x, y, z, out, cpuTimeTotal=tTotal, cpuTimeSetup=tSetup, numProcs=world)
            if verbose:
                print(" Wrote netcdf results")
    return stats


def main(argv=None):
    """python -m i3rc_monte_carlo_model_b200.driver namelist.nml   (under torchrun: one rank per GPU)"""
    import os
    import sys
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("usage: python -m i3rc_monte_carlo_model_b200.driver <namelist file>", file=sys.stderr)
        return 2
    dist = None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from ._lib import backend
    be = backend()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if be.set_device(local) != 0:
        print("monteCarloDriver: no CUDA device (the integrator has no CPU fallback)", file=sys.stderr)
        return 1
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        monteCarloDriver(argv[0], backend=be, dist=dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
