#!/usr/bin/env python
"""Writes an I3RC domain file + a monteCarloDriver namelist into a directory (for trying the drivers by hand):
make_driver_case.py <dir> [landsat|step] [photons per batch] [batches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i3rc_monte_carlo_model_b200 import fields, fileIO

out = sys.argv[1]
case = sys.argv[2] if len(sys.argv) > 2 else "step"
nph = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 8
os.makedirs(out, exist_ok=True)
dom = os.path.join(out, case + ".dom")
fileIO.write_Domain(fields.landsat_cloud(1.0) if case == "landsat" else fields.step_cloud(0.99), dom)
open(os.path.join(out, "run.nml"), "w").write(f"""&radiativeTransfer
  solarFlux = 1., solarMu = .5, solarAzimuth = 0., surfaceAlbedo = 0.,
  intensityMus  = 1., .5, .5, intensityPhis = 0., 0., 180.
/
&monteCarlo
  numPhotonsPerBatch = {nph}, numBatches = {nb}, iseed = 10, nPhaseIntervals = 10001
/
&algorithms
  useRayTracing = .true., useRussianRoulette = .true., useRussianRouletteForIntensity = .true., zetaMin = 0.3
/
&output
  reportAbsorptionProfile = .true., reportVolumeAbsorption = .false.
/
&fileNames
  domainFileName = "{dom}",
  outputFluxFile = "{out}/flux.txt", outputRadFile = "{out}/rad.txt", outputNetcdfFile = "{out}/results.nc"
/
""")
print(os.path.join(out, "run.nml"))
