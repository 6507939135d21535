// planeParallel for the B200 integrator: Example-Drivers/planeParallel.f95 written against the C ABI.  Builds a
// homogeneous slab in memory (createDomain, :299-379: Henyey-Greenstein by Legendre moments or angle/value pairs, or
// an entry of a phase-function-table file), runs numBatches batches seeded (/ batch, iseed /) (:202-236) and prints
// the mean and the batch standard deviation of fluxes or radiances in the reference's formats (:241-273).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/i3rc_b200.h"
#include "domain_io.hpp"
#include "namelist.hpp"

using namespace i3rc_host;

static void check(int rc, i3rc_integrator* h, const char* what) {
  if (rc == I3RC_SUCCESS) return;
  fprintf(stderr, " %s: %s\n", what, i3rc_last_message(h));
  if (rc == I3RC_FAILURE) exit(1);
}

int main(int argc, char** argv) {
  if (argc != 2) {
    fprintf(stderr, "usage: planeParallel <namelist file>\n");
    return 2;
  }
  Namelist nml;
  if (!nml.load(argv[1])) {
    fprintf(stderr, "planeParallel: can't read namelist file %s\n", argv[1]);
    return 1;
  }
  const float solarMu = (float)nml.real("radiativeTransfer", "solarMu", 0.5), solarAzimuth = (float)nml.real("radiativeTransfer", "solarAzimuth", 0.0);
  const float surfaceAlbedo = (float)nml.real("radiativeTransfer", "surfaceAlbedo", 0.0);
  std::vector<float> mus, phis;
  for (double v : nml.reals("radiativeTransfer", "intensityMus")) mus.push_back((float)v);
  for (double v : nml.reals("radiativeTransfer", "intensityPhis")) phis.push_back((float)v);
  int nAng = 0;
  for (float v : mus) nAng += std::fabs(v) > 0.0f;
  mus.resize(nAng);
  phis.resize(nAng, 0.0f);
  const bool computeIntensity = nAng > 0;
  const long numPhotonsPerBatch = nml.integer("monteCarlo", "numPhotonsPerBatch", 100000);
  const int numBatches = (int)nml.integer("monteCarlo", "numBatches", 4), iseed = (int)nml.integer("monteCarlo", "iseed", 10);
  const bool useRayTracing = nml.logical("algorithms", "useRayTracing", true), useRussianRoulette = nml.logical("algorithms", "useRussianRoulette", true);
  const bool useHybrid = nml.logical("algorithms", "useHybridPhaseFunsForIntenCalcs", false);
  const float hybridWidth = (float)nml.real("algorithms", "hybridPhaseFunWidth", 7.0);
  const bool useRRI = nml.logical("algorithms", "useRussianRouletteForIntensity", true);
  const float zetaMin = (float)nml.real("algorithms", "zetaMin", 0.0);
  const float SSA = (float)nml.real("problemOptics", "SSA", 1.0), opticalDepth = (float)nml.real("problemOptics", "opticalDepth", 1.0);
  const float g = (float)nml.real("problemOptics", "g", 0.85);
  const int nLegendre = (int)nml.integer("problemOptics", "nLegendreCoefficients", 64), nAngles = (int)nml.integer("problemOptics", "nAngles", 5000);
  const bool useMoments = nml.logical("problemOptics", "useMoments", true);
  const std::string tableFile = nml.str("problemOptics", "phaseFunctionTableFile", "");
  const int tableIndex = (int)nml.integer("problemOptics", "phaseFunctionTableIndex", 0);
  const float domainSize = (float)nml.real("problemDomain", "domainSize", 500.0), thickness = (float)nml.real("problemDomain", "physicalThickness", 250.0);
  const int nLayers = (int)nml.integer("problemDomain", "nLayers", 1), nX = (int)nml.integer("problemDomain", "nx", 1), nY = (int)nml.integer("problemDomain", "ny", 1);
  const bool useSurfaceProperties = nml.logical("problemDomain", "useSurfaceProperties", false);
  const std::string domainFileName = nml.str("filenames", "domainFileName", "");

  // ---- createDomain (:299-379)
  Domain d;
  for (int i = 0; i <= nX; i++) d.x.push_back(domainSize / nX * (float)i);
  for (int i = 0; i <= nY; i++) d.y.push_back(domainSize / nY * (float)i);
  for (int i = 0; i <= nLayers; i++) d.z.push_back(thickness / (float)nLayers * (float)i);
  Component cloud;
  cloud.name = "cloud";
  cloud.nz = nLayers;
  const size_t ncell = (size_t)nX * nY * nLayers;
  int pfIndex = 1;
  if (!tableFile.empty()) {
    NcFile f;
    std::string err;
    if (!f.open(tableFile) || !read_phase_table(f, "", cloud.table, err)) {
      fprintf(stderr, " read_PhaseFunctionTable: can't read %s %s\n", tableFile.c_str(), err.c_str());
      return 1;
    }
    pfIndex = tableIndex;
  } else if (useMoments) {
    cloud.table.kind = 1;
    cloud.table.key = {1.0f};
    cloud.table.offsets = {0, nLegendre};
    for (int i = 1; i <= nLegendre; i++) cloud.table.coefs.push_back(std::pow(g, (float)i));
  } else {
    cloud.table.kind = 2;
    cloud.table.key = {1.0f};
    const float pi = std::acos(-1.0f);
    for (int i = 0; i < nAngles; i++) {
      float a = (float)i / (float)(nAngles - 1) * pi;
      cloud.table.angles.push_back(a);
      cloud.table.values.push_back((1 - g * g) / std::pow(1 + g * g - 2 * g * std::cos(a), 1.5f));
    }
  }
  cloud.ext.assign(ncell, opticalDepth / thickness);
  cloud.ssa.assign(ncell, SSA);
  cloud.pfi.assign(ncell, pfIndex);
  d.comps.push_back(cloud);
  if (!domainFileName.empty()) {
    if (!write_domain(d, domainFileName)) {
      fprintf(stderr, " write_Domain: error writing %s\n", domainFileName.c_str());
      return 1;
    }
    printf(" Wrote domain to file %s\n", domainFileName.c_str());
  }

  if (i3rc_device_count() <= 0 || i3rc_set_device(0) != I3RC_SUCCESS) {
    fprintf(stderr, "planeParallel: no CUDA device (the integrator has no CPU fallback)\n");
    return 1;
  }
  i3rc_component comp;
  memset(&comp, 0, sizeof comp);
  comp.extinction = d.comps[0].ext.data();
  comp.ssa = d.comps[0].ssa.data();
  comp.phase_index = d.comps[0].pfi.data();
  comp.z_level_base = 1;
  comp.nz = nLayers;
  comp.table = d.comps[0].table.as_c();
  i3rc_integrator* h = nullptr;
  if (i3rc_new_Integrator_components(nX, nY, nLayers, d.x.data(), d.y.data(), d.z.data(), 1, &comp, &h) == I3RC_FAILURE) {
    fprintf(stderr, " new_Integrator: %s\n", i3rc_last_message(nullptr));
    return 1;
  }
  i3rc_params p;
  memset(&p, 0, sizeof p);
  const float one = 0.0f, edges[2] = {0.0f, 1.0f};
  (void)one;
  if (useSurfaceProperties) {  // a surfaceDescription with one Lambertian albedo (:145-151)
    p.present = I3RC_P_surfaceBDRF;
    p.surf_nx = p.surf_ny = 1;
    p.surf_x = edges;
    p.surf_y = edges;
    p.surf_params = &surfaceAlbedo;
  } else {
    p.present = I3RC_P_surfaceAlbedo;
    p.surfaceAlbedo = surfaceAlbedo;
  }
  check(i3rc_specifyParameters(h, &p), h, "specifyParameters");
  if (computeIntensity) {
    memset(&p, 0, sizeof p);
    p.present = I3RC_P_intensityMus | I3RC_P_intensityPhis;
    p.numIntensityDirections = nAng;
    p.intensityMus = mus.data();
    p.intensityPhis = phis.data();
    check(i3rc_specifyParameters(h, &p), h, "specifyParameters");
  }
  memset(&p, 0, sizeof p);
  p.present = I3RC_P_useRayTracing | I3RC_P_useRussianRoulette | I3RC_P_useHybridPhaseFunsForIntenCalcs | I3RC_P_hybridPhaseFunWidth |
              I3RC_P_useRussianRouletteForIntensity | I3RC_P_zetaMin;
  p.useRayTracing = useRayTracing;
  p.useRussianRoulette = useRussianRoulette;
  p.useHybridPhaseFunsForIntenCalcs = useHybrid;
  p.hybridPhaseFunWidth = hybridWidth;
  p.useRussianRouletteForIntensity = useRRI;
  p.zetaMin = zetaMin;
  check(i3rc_specifyParameters(h, &p), h, "specifyParameters");

  if (numBatches > 0 && numPhotonsPerBatch > 0) {
    const size_t ncol = (size_t)nX * nY;
    std::vector<float> fluxUp(ncol * numBatches), fluxDown(ncol * numBatches), fluxAbs(ncol * numBatches), inten;
    if (computeIntensity) inten.resize(ncol * nAng * numBatches);
    i3rc_photon_source src;
    memset(&src, 0, sizeof src);
    src.kind = I3RC_SRC_DIRECTIONAL;
    src.solarMu = solarMu;
    src.solarAzimuth = solarAzimuth;
    src.numberOfPhotons = numPhotonsPerBatch;
    for (int batch = 1; batch <= numBatches; batch++) {
      const int32_t seed[2] = {batch, iseed};  // (/ batch, iseed /), :207
      check(i3rc_computeRadiativeTransfer(h, &src, seed, 2), h, "computeRadiativeTransfer");
      float mUp, mDown, mAbs;
      if (computeIntensity)
        check(i3rc_reportResults(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                 inten.data() + (size_t)(batch - 1) * ncol * nAng), h, "reportResults");
      else
        check(i3rc_reportResults(h, &mUp, &mDown, &mAbs, fluxUp.data() + (size_t)(batch - 1) * ncol, fluxDown.data() + (size_t)(batch - 1) * ncol,
                                 fluxAbs.data() + (size_t)(batch - 1) * ncol, nullptr, nullptr, nullptr, nullptr), h, "reportResults");
    }
    const double theta0 = std::acos(solarMu) * 180.0 / std::acos(-1.0);
    if (computeIntensity) {
      printf("   tau  omega   g  theta0    mu   phi radiance    error\n");
      for (int i = 0; i < nAng; i++) {
        double mean = 0;
        std::vector<double> perBatch(numBatches, 0.0);
        for (int b = 0; b < numBatches; b++) {
          for (size_t k = 0; k < ncol; k++) perBatch[b] += inten[((size_t)b * nAng + i) * ncol + k];
          mean += perBatch[b];
          perBatch[b] /= (double)ncol;
        }
        mean /= (double)numBatches * ncol;
        double var = 0;
        for (int b = 0; b < numBatches; b++) var += (perBatch[b] - mean) * (perBatch[b] - mean);
        printf("%6.2f %5.3f %5.3f  %5.2f %7.5f %3d %8.6f %10.8f\n", opticalDepth, SSA, g, theta0, mus[i], (int)phis[i], mean,
               std::sqrt(var / numBatches));
      }
    } else {
      auto stats = [&](const std::vector<float>& a, double& mean, double& sd) {
        mean = 0;
        for (float v : a) mean += v;
        mean /= (double)a.size();
        sd = 0;
        if (numBatches > 1) {
          for (float v : a) sd += (v - mean) * (v - mean);
          sd = std::sqrt(sd / ((double)(numBatches - 1) * ncol));
        }
      };
      double mu_, sdu, md, sdd, ma, sda;
      stats(fluxUp, mu_, sdu);
      stats(fluxDown, md, sdd);
      stats(fluxAbs, ma, sda);
      printf("   tau  omega   g  theta0   Fup      Fdn    FluxUpErr FluxDownErr FluxAbs FluxAbsErr\n");
      printf("%6.2f %5.3f %5.3f  %5.2f %7.5f   %7.5f   %7.5f   %7.5f   %7.5f   %7.5f   \n", opticalDepth, SSA, g, theta0, mu_, md, sdu, sdd, ma, sda);
    }
  }
  i3rc_finalize_Integrator(h);
  return 0;
}
