// monteCarloDriver for the B200 integrator: the program of Example-Drivers/monteCarloDriver.f95 written against the
// C ABI (include/i3rc_b200.h) -- the compiled host side of the drop-in while no Fortran compiler is at hand.
//
//   monteCarloDriver run.nml            one process per GPU; ranks come from the launcher's environment
//                                       (RANK / WORLD_SIZE / LOCAL_RANK as torchrun sets them, or OMPI_COMM_WORLD_*)
//
// Flow, line by line after the reference: five namelists (:143-150) -> read_Domain (:159) -> new_Integrator (:169)
// -> four specifyParameters calls (:174-216) -> one photon with seed (/iseed,0/) (:240-253) -> this rank's block of
// batches (:264-326, on the device: i3rc_run_batches) -> ONE all-reduce of the moment buffer (replaces :333-348)
// -> mean / standard error (:358-378) -> rank 0 writes the ASCII and netCDF result files (:382-419).
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/i3rc_b200.h"
#include "domain_io.hpp"
#include "namelist.hpp"
#include "results_io.hpp"

using namespace i3rc_host;

static int env_int(const char* a, const char* b, int dflt) {
  const char* v = getenv(a);
  if (!v && b) v = getenv(b);
  return v ? atoi(v) : dflt;
}

// printStatus (Code/userInterface_Unix.f95:21-54): print warnings, stop on failure
static void check(int rc, i3rc_integrator* h, const char* what) {
  if (rc == I3RC_SUCCESS) return;
  fprintf(stderr, " %s: %s\n", what, i3rc_last_message(h));
  if (rc == I3RC_FAILURE) exit(1);
}

int main(int argc, char** argv) {
  if (argc != 2) {
    fprintf(stderr, "usage: monteCarloDriver <namelist file>\n");
    return 2;
  }
  const auto t0 = std::chrono::steady_clock::now();
  auto seconds = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
  const int thisProc = env_int("RANK", "OMPI_COMM_WORLD_RANK", 0), numProcs = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1);
  const int local = env_int("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", 0);
  const bool master = thisProc == 0;

  Namelist nml;
  if (!nml.load(argv[1])) {
    fprintf(stderr, "monteCarloDriver: can't read namelist file %s\n", argv[1]);
    return 1;
  }
  RunConfig c;
  c.solarFlux = nml.real("radiativeTransfer", "solarFlux", 1.0);
  c.solarMu = nml.real("radiativeTransfer", "solarMu", 1.0);
  c.solarAzimuth = nml.real("radiativeTransfer", "solarAzimuth", 0.0);
  c.surfaceAlbedo = nml.real("radiativeTransfer", "surfaceAlbedo", 0.0);
  for (double v : nml.reals("radiativeTransfer", "intensityMus")) c.intensityMus.push_back((float)v);
  for (double v : nml.reals("radiativeTransfer", "intensityPhis")) c.intensityPhis.push_back((float)v);
  int numRadDir = 0;  // count(abs(intensityMus) > 0), :151
  for (float v : c.intensityMus) numRadDir += std::fabs(v) > 0.0f;
  c.intensityMus.resize(numRadDir);
  c.intensityPhis.resize(numRadDir, 0.0f);
  c.numPhotonsPerBatch = nml.integer("monteCarlo", "numPhotonsPerBatch", 0);
  c.numBatches = (int)nml.integer("monteCarlo", "numBatches", 100);
  c.iseed = (int)nml.integer("monteCarlo", "iseed", 10);
  c.nPhaseIntervals = (int)nml.integer("monteCarlo", "nPhaseIntervals", 10001);
  c.useRayTracing = nml.logical("algorithms", "useRayTracing", true);
  c.useRussianRoulette = nml.logical("algorithms", "useRussianRoulette", true);
  c.useHybridPhaseFunsForIntenCalcs = nml.logical("algorithms", "useHybridPhaseFunsForIntenCalcs", false);
  c.hybridPhaseFunWidth = nml.real("algorithms", "hybridPhaseFunWidth", 7.0);
  c.numOrdersOrigPhaseFunIntenCalcs = (int)nml.integer("algorithms", "numOrdersOrigPhaseFunIntenCalcs", 0);
  c.useRussianRouletteForIntensity = nml.logical("algorithms", "useRussianRouletteForIntensity", true);
  c.zetaMin = nml.real("algorithms", "zetaMin", 0.3);
  c.limitIntensityContributions = nml.logical("algorithms", "limitIntensityContributions", false);
  c.maxIntensityContribution = nml.real("algorithms", "maxIntensityContribution", 77.0);
  c.reportVolumeAbsorption = nml.logical("output", "reportVolumeAbsorption", false);
  c.reportAbsorptionProfile = nml.logical("output", "reportAbsorptionProfile", false);
  c.domainFileName = nml.str("fileNames", "domainFileName", "");
  c.outputFluxFile = nml.str("fileNames", "outputFluxFile", "");
  c.outputRadFile = nml.str("fileNames", "outputRadFile", "");
  c.outputAbsProfFile = nml.str("fileNames", "outputAbsProfFile", "");
  c.outputAbsVolumeFile = nml.str("fileNames", "outputAbsVolumeFile", "");
  c.outputNetcdfFile = nml.str("fileNames", "outputNetcdfFile", "");
  const bool computeIntensity = numRadDir > 0 && (!c.outputRadFile.empty() || !c.outputNetcdfFile.empty());
  if (!computeIntensity) c.outputRadFile.clear();

  if (i3rc_device_count() <= 0 || i3rc_set_device(local) != I3RC_SUCCESS) {
    fprintf(stderr, "monteCarloDriver: no CUDA device (the integrator has no CPU fallback)\n");
    return 1;
  }
  Domain dom;
  std::string err;
  if (!read_domain(c.domainFileName, dom, err)) {
    fprintf(stderr, " %s\n", err.c_str());
    return 1;
  }
  const int nx = dom.nx(), ny = dom.ny(), nz = dom.nz();
  std::vector<i3rc_component> comps(dom.comps.size());
  for (size_t i = 0; i < comps.size(); i++) {
    const Component& k = dom.comps[i];
    comps[i].extinction = k.ext.data();
    comps[i].ssa = k.ssa.data();
    comps[i].phase_index = k.pfi.data();
    comps[i].horizontally_uniform = k.uniform;
    comps[i].z_level_base = k.zLevelBase;
    comps[i].nz = k.nz;
    comps[i].table = k.table.as_c();
  }
  i3rc_integrator* h = nullptr;
  if (i3rc_new_Integrator_components(nx, ny, nz, dom.x.data(), dom.y.data(), dom.z.data(), (int)comps.size(), comps.data(), &h) ==
      I3RC_FAILURE) {
    fprintf(stderr, " new_Integrator: %s\n", i3rc_last_message(nullptr));
    return 1;
  }
  const std::vector<float> x = dom.x, y = dom.y, z = dom.z;
  dom.comps.clear();  // finalize_Domain: the integrator holds its own copies (:167-171)

  i3rc_params p;
  memset(&p, 0, sizeof p);
  p.present = I3RC_P_surfaceAlbedo | I3RC_P_minInverseTableSize;
  p.surfaceAlbedo = (float)c.surfaceAlbedo;
  p.minInverseTableSize = c.nPhaseIntervals;
  check(i3rc_specifyParameters(h, &p), h, "specifyParameters");
  if (computeIntensity) {
    memset(&p, 0, sizeof p);
    p.present = I3RC_P_minForwardTableSize | I3RC_P_intensityMus | I3RC_P_intensityPhis | I3RC_P_computeIntensity;
    p.minForwardTableSize = c.nPhaseIntervals;
    p.numIntensityDirections = numRadDir;
    p.intensityMus = c.intensityMus.data();
    p.intensityPhis = c.intensityPhis.data();
    p.computeIntensity = 1;
    check(i3rc_specifyParameters(h, &p), h, "specifyParameters");
  }
  memset(&p, 0, sizeof p);
  p.present = I3RC_P_useRayTracing | I3RC_P_useRussianRoulette;
  p.useRayTracing = c.useRayTracing;
  p.useRussianRoulette = c.useRussianRoulette;
  check(i3rc_specifyParameters(h, &p), h, "specifyParameters");
  if (computeIntensity) {
    memset(&p, 0, sizeof p);
    p.present = I3RC_P_useHybridPhaseFunsForIntenCalcs | I3RC_P_hybridPhaseFunWidth | I3RC_P_numOrdersOrigPhaseFunIntenCalcs |
                I3RC_P_useRussianRouletteForIntensity | I3RC_P_zetaMin | I3RC_P_limitIntensityContributions |
                I3RC_P_maxIntensityContribution;
    p.useHybridPhaseFunsForIntenCalcs = c.useHybridPhaseFunsForIntenCalcs;
    p.hybridPhaseFunWidth = (float)c.hybridPhaseFunWidth;
    p.numOrdersOrigPhaseFunIntenCalcs = c.numOrdersOrigPhaseFunIntenCalcs;
    p.useRussianRouletteForIntensity = c.useRussianRouletteForIntensity;
    p.zetaMin = (float)c.zetaMin;
    p.limitIntensityContributions = c.limitIntensityContributions;
    p.maxIntensityContribution = (float)c.maxIntensityContribution;
    check(i3rc_specifyParameters(h, &p), h, "specifyParameters");
  }

  // one photon, seed (/ iseed, 0 /): checks the set-up and builds the tables before the clock of the batches starts
  i3rc_photon_source src;
  memset(&src, 0, sizeof src);
  src.kind = I3RC_SRC_DIRECTIONAL;
  src.solarMu = (float)c.solarMu;
  src.solarAzimuth = (float)c.solarAzimuth;
  src.numberOfPhotons = 1;
  if (!i3rc_isReady_Integrator(h)) {
    fprintf(stderr, "Integrator is not ready.\n");
    return 1;
  }
  const int32_t seed0[2] = {c.iseed, 0};
  check(i3rc_computeRadiativeTransfer(h, &src, seed0, 2), h, "computeRadiativeTransfer");
  const double cpuSetup = seconds();
  if (master) printf(" Setup CPU time (secs, approx): %d\n", (int)cpuSetup);

  // batches per process (:264-274)
  c.numBatches = c.numBatches < 2 ? 2 : c.numBatches;
  int bpp = c.numBatches / numProcs;
  if (c.numBatches % numProcs != 0) {
    bpp++;
    c.numBatches = bpp * numProcs;
  }
  if (master) printf(" Doing %d batches on each of %d processors.\n", bpp, numProcs);
  const bool withVolume = c.reportVolumeAbsorption || !c.outputAbsVolumeFile.empty();
  check(i3rc_stats_reset(h, withVolume), h, "stats_reset");
  src.numberOfPhotons = c.numPhotonsPerBatch;
  check(i3rc_run_batches(h, &src, c.iseed, 0, thisProc * bpp + 1, bpp), h, "computeRadiativeTransfer");

  if (numProcs > 1) {  // the one collective: NCCL all-reduce of the packed moment buffer
    // The 128-byte NCCL id travels through a file next to the namelist (no MPI at hand); rank 0 writes, the others poll.
    const char* rv = getenv("I3RC_RENDEZVOUS_FILE");
    std::string path = rv ? rv : std::string(argv[1]) + ".ncclid." + std::to_string(env_int("MASTER_PORT", nullptr, 0));
    char id[128];
    if (master) {
      check(i3rc_comm_unique_id(id), h, "comm_unique_id");
      std::string tmp = path + ".tmp";
      FILE* f = fopen(tmp.c_str(), "wb");
      if (!f || fwrite(id, 1, 128, f) != 128) {
        fprintf(stderr, "monteCarloDriver: can't write %s\n", tmp.c_str());
        return 1;
      }
      fclose(f);
      rename(tmp.c_str(), path.c_str());
    } else {
      for (int tries = 0;; tries++) {
        FILE* f = fopen(path.c_str(), "rb");
        if (f) {
          size_t n = fread(id, 1, 128, f);
          fclose(f);
          if (n == 128) break;
        }
        if (tries > 6000) {
          fprintf(stderr, "monteCarloDriver: rank %d never saw %s\n", thisProc, path.c_str());
          return 1;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
      }
    }
    check(i3rc_comm_init(h, numProcs, thisProc, id), h, "comm_init");
    check(i3rc_stats_allreduce(h), h, "stats_allreduce");
    check(i3rc_comm_finalize(h), h, "comm_finalize");
    if (master) remove(path.c_str());
  }

  const size_t ncol = (size_t)nx * ny, ncell = ncol * nz;
  Stats s;
  s.fluxUp.resize(2 * ncol);
  s.fluxDown.resize(2 * ncol);
  s.fluxAbsorbed.resize(2 * ncol);
  s.absorbedProfile.resize(2 * (size_t)nz);
  if (withVolume) s.absorbedVolume.resize(2 * ncell);
  if (computeIntensity) {
    s.radiance.resize(2 * (size_t)numRadDir * ncol);
    s.meanRadiance.resize(2 * (size_t)numRadDir);
  }
  i3rc_stats_out o;
  memset(&o, 0, sizeof o);
  o.meanFluxUp = s.meanFluxUp;
  o.meanFluxDown = s.meanFluxDown;
  o.meanFluxAbsorbed = s.meanFluxAbsorbed;
  o.fluxUp = s.fluxUp.data();
  o.fluxDown = s.fluxDown.data();
  o.fluxAbsorbed = s.fluxAbsorbed.data();
  o.absorbedProfile = s.absorbedProfile.data();
  o.absorbedVolume = withVolume ? s.absorbedVolume.data() : nullptr;
  o.radiance = computeIntensity ? s.radiance.data() : nullptr;
  o.meanRadiance = computeIntensity ? s.meanRadiance.data() : nullptr;
  check(i3rc_stats_report(h, c.solarFlux, c.numBatches, &o), h, "stats_report");
  const double cpuTotal = seconds();
  if (master) {
    printf(" Total CPU time (secs, approx): %d\n", (int)cpuTotal);
    if (!c.outputFluxFile.empty() || !c.outputAbsProfFile.empty() || !c.outputAbsVolumeFile.empty() || !c.outputRadFile.empty()) {
      write_results_ascii(c, x, y, z, s);
      printf(" Wrote ASCII results\n");
    }
    if (!c.outputNetcdfFile.empty()) {
      if (!write_results_netcdf(c, x, y, z, s, computeIntensity, cpuTotal, cpuSetup, numProcs)) {
        fprintf(stderr, "monteCarloDriver: can't write %s\n", c.outputNetcdfFile.c_str());
        return 1;
      }
      printf(" Wrote netcdf results\n");
    }
  }
  i3rc_finalize_Integrator(h);
  return 0;
}
