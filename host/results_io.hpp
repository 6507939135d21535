// Result files of Example-Drivers/monteCarloDriver.f95: writeResults_ASCII (:436-605) and writeResults_netcdf
// (:609-854), with the reference's edit descriptors, names, attributes and dimension order.
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "netcdf3.hpp"

namespace i3rc_host {

struct RunConfig {  // the five namelists of monteCarloDriver.f95:61-103
  double solarFlux = 1.0, solarMu = 1.0, solarAzimuth = 0.0, surfaceAlbedo = 0.0;
  std::vector<float> intensityMus, intensityPhis;
  long numPhotonsPerBatch = 0;
  int numBatches = 100, iseed = 10, nPhaseIntervals = 10001;
  bool useRayTracing = true, useRussianRoulette = true, useHybridPhaseFunsForIntenCalcs = false;
  double hybridPhaseFunWidth = 7.0;
  int numOrdersOrigPhaseFunIntenCalcs = 0;
  bool useRussianRouletteForIntensity = true;
  double zetaMin = 0.3;
  bool limitIntensityContributions = false;
  double maxIntensityContribution = 77.0;
  bool reportVolumeAbsorption = false, reportAbsorptionProfile = false;
  std::string domainFileName, outputFluxFile, outputRadFile, outputAbsProfFile, outputAbsVolumeFile, outputNetcdfFile;
};

// mean and standard error of every output, [2][...] like i3rc_stats_out (C order: [z][y][x], [dir][y][x])
struct Stats {
  double meanFluxUp[2], meanFluxDown[2], meanFluxAbsorbed[2];
  std::vector<double> fluxUp, fluxDown, fluxAbsorbed, absorbedProfile, absorbedVolume, radiance, meanRadiance;
};

inline std::string F(double v, int w, int d) {  // Fortran Fw.d
  char b[64];
  snprintf(b, sizeof b, "%*.*f", w, d, v);
  std::string s(b);
  if ((int)s.size() > w && s.compare(0, 2, "0.") == 0) s.erase(0, 1);  // the optional leading zero goes first
  if ((int)s.size() > w && s.compare(0, 3, "-0.") == 0) s.erase(1, 1);
  return (int)s.size() <= w ? s : std::string(w, '*');
}
inline std::string E13_6(double v) {  // Fortran E13.6: 0.ddddddE+ee
  if (v == 0) return " 0.000000E+00";
  int e = (int)std::floor(std::log10(std::fabs(v))) + 1;
  double m = v / std::pow(10.0, e);
  if (std::fabs(std::round(m * 1e6) / 1e6) >= 1.0) {
    m /= 10.0;
    e++;
  }
  char b[64];
  snprintf(b, sizeof b, "%9.6fE%+03d", m, e);
  std::string s(b);
  return s.size() < 13 ? std::string(13 - s.size(), ' ') + s : s;
}
inline const char* L(bool b) { return b ? "T" : "F"; }
inline std::string pair_(double m, double s) { return " " + F(m, 9, 4) + " " + F(s, 9, 4); }

inline void header(FILE* fh, const char* title, const RunConfig& c, const char* outType, bool radiance = false) {
  fprintf(fh, "!   I3RC Monte Carlo 3D Solar Radiative Transfer: %s\n", title);
  fprintf(fh, "!  Property_File=%-60.60s\n", c.domainFileName.c_str());  // A60 of a blank-padded character(256)
  fprintf(fh, "!  Num_Photons=%10ld\n", c.numPhotonsPerBatch * c.numBatches);
  fprintf(fh, "!  PhotonTracing=%s    Russian_Roulette=%s\n", L(c.useRayTracing), L(c.useRussianRoulette));
  fprintf(fh, "!  Hybrid_Phase_Func_for_Radiance=%s   Gaussian_Phase_Func_Width_deg=%s\n", L(c.useHybridPhaseFunsForIntenCalcs),
          F(c.hybridPhaseFunWidth, 5, 2).c_str());
  if (radiance) {
    fprintf(fh, "!  Intensity_uses_Russian_Roulette=%s   Intensity_Russian_Roulette_zeta_min=%s\n", L(c.useRussianRouletteForIntensity),
            F(c.zetaMin, 5, 2).c_str());
    fprintf(fh, "!  limited_intensity_contributions=%s   max_intensity_contribution=%s\n", L(c.limitIntensityContributions),
            F(c.maxIntensityContribution, 5, 2).c_str());
  }
  fprintf(fh, "!  Solar_Flux=%s   Solar_Mu=%s   Solar_Phi=%s\n", E13_6(c.solarFlux).c_str(), F(c.solarMu, 10, 7).c_str(),
          F(c.solarAzimuth, 7, 3).c_str());
  fprintf(fh, "!  Lambertian_Surface_Albedo=%s\n", F(c.surfaceAlbedo, 7, 4).c_str());
  fprintf(fh, "!  Output_Type= %s\n", outType);
}

inline void write_results_ascii(const RunConfig& c, const std::vector<float>& x, const std::vector<float>& y, const std::vector<float>& z,
                                const Stats& s) {
  const int nx = (int)x.size() - 1, ny = (int)y.size() - 1, nz = (int)z.size() - 1;
  const size_t ncol = (size_t)nx * ny;
  auto mid = [](const std::vector<float>& e, int i) { return ((double)e[i] + (double)e[i + 1]) / 2.0; };
  if (!c.outputFluxFile.empty()) {
    FILE* fh = fopen(c.outputFluxFile.c_str(), "w");
    if (fh) {
      header(fh, "Flux", c, "Pixel Flux");
      fprintf(fh, "!  Upwelling_Level=%s   Downwelling_level=%s\n", F(z[nz], 7, 3).c_str(), F(z[0], 7, 3).c_str());
      fprintf(fh, "!   X      Y           Flux_Up             Flux_Down            Flux_Absorbed \n");
      fprintf(fh, "!                  Mean     StdErr       Mean     StdErr       Mean     StdErr\n");
      fprintf(fh, "%14s %s %s %s\n", "!  Average:   ", pair_(s.meanFluxUp[0], s.meanFluxUp[1]).c_str(),
              pair_(s.meanFluxDown[0], s.meanFluxDown[1]).c_str(), pair_(s.meanFluxAbsorbed[0], s.meanFluxAbsorbed[1]).c_str());
      for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
          size_t k = (size_t)j * nx + i;
          fprintf(fh, "%s%s %s %s %s\n", F(mid(x, i), 7, 3).c_str(), F(mid(y, j), 7, 3).c_str(), pair_(s.fluxUp[k], s.fluxUp[ncol + k]).c_str(),
                  pair_(s.fluxDown[k], s.fluxDown[ncol + k]).c_str(), pair_(s.fluxAbsorbed[k], s.fluxAbsorbed[ncol + k]).c_str());
        }
      fclose(fh);
    }
  }
  if (!c.outputAbsProfFile.empty()) {
    FILE* fh = fopen(c.outputAbsProfFile.c_str(), "w");
    if (fh) {
      header(fh, "Absorption Profile", c, "Absorption Profile");
      fprintf(fh, "!   Z    Absorbed_Flux (flux/km) \n!          Mean     StdErr \n");
      for (int k = 0; k < nz; k++) fprintf(fh, "%s %s\n", F(mid(z, k), 7, 3).c_str(), pair_(s.absorbedProfile[k], s.absorbedProfile[nz + k]).c_str());
      fclose(fh);
    }
  }
  if (!c.outputAbsVolumeFile.empty() && !s.absorbedVolume.empty()) {
    FILE* fh = fopen(c.outputAbsVolumeFile.c_str(), "w");
    if (fh) {
      header(fh, "3D Absorption Field", c, "Volume Absorption ");
      fprintf(fh, "!    X       Y        Z       Absorbed_Flux (flux/km)\n!                               Mean     StdErr \n");
      const size_t ncell = ncol * nz;
      for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++)
          for (int k = 0; k < nz; k++) {
            size_t o = ((size_t)k * ny + j) * nx + i;
            fprintf(fh, "%s %s %s %s\n", F(mid(x, i), 7, 3).c_str(), F(mid(y, j), 7, 3).c_str(), F(mid(z, k), 7, 3).c_str(),
                    pair_(s.absorbedVolume[o], s.absorbedVolume[ncell + o]).c_str());
          }
      fclose(fh);
    }
  }
  if (!c.outputRadFile.empty() && !s.radiance.empty()) {
    const int nd = (int)c.intensityMus.size();
    FILE* fh = fopen(c.outputRadFile.c_str(), "w");
    if (fh) {
      header(fh, "Radiance", c, "Pixel Radiance", true);
      fprintf(fh, "!  RADIANCE AT Z=%s   NXO=%4d   NYO=%4d   NDIR=%4d\n", F(z[nz], 7, 3).c_str(), nx, ny, nd);
      fprintf(fh, "!   X      Y         Radiance (Mean, StdErr)\n");
      for (int k = 0; k < nd; k++) {
        fprintf(fh, "!  %s %s  <- (mu,phi)\n", F(c.intensityMus[k], 8, 5).c_str(), F(c.intensityPhis[k], 6, 2).c_str());
        for (int j = 0; j < ny; j++)
          for (int i = 0; i < nx; i++) {
            size_t o = ((size_t)k * ny + j) * nx + i;
            fprintf(fh, "%s%s%s\n", F(mid(x, i), 7, 3).c_str(), F(mid(y, j), 7, 3).c_str(), pair_(s.radiance[o], s.radiance[(size_t)nd * ncol + o]).c_str());
          }
      }
      fclose(fh);
    }
  }
}

inline bool write_results_netcdf(const RunConfig& c, const std::vector<float>& x, const std::vector<float>& y, const std::vector<float>& z,
                                 const Stats& s, bool computeIntensity, double cpuTotal, double cpuSetup, int numProcs) {
  const int nx = (int)x.size() - 1, ny = (int)y.size() - 1, nz = (int)z.size() - 1;
  const size_t ncol = (size_t)nx * ny, ncell = ncol * nz;
  NcFile f;
  f.put_att_text("description", "Output from I3RC Community Monte Carlo Model");
  f.put_att_text("Domain_filename", c.domainFileName);
  f.put_att("Surface_albedo", NC_FLOAT, c.surfaceAlbedo);
  f.put_att("Total_number_of_photons", NC_INT, c.numPhotonsPerBatch * c.numBatches);
  f.put_att("Number_of_batches", NC_INT, c.numBatches);
  f.put_att("Solar_flux", NC_FLOAT, c.solarFlux);
  f.put_att("Solar_mu", NC_FLOAT, c.solarMu);
  f.put_att("Solar_phi", NC_FLOAT, c.solarAzimuth);
  f.put_att("Random_number_seed", NC_INT, c.iseed);
  f.put_att("Phase_function_table_sizes", NC_INT, c.nPhaseIntervals);
  f.put_att_text("Algorithm", c.useRayTracing ? "Ray_tracing" : "Max_cross_section");
  f.put_att("Intensity_uses_hyrbid_phase_functions", NC_INT, (int)c.useHybridPhaseFunsForIntenCalcs);  // (sic)
  f.put_att("Hybrid_phase_function_width", NC_FLOAT, c.useHybridPhaseFunsForIntenCalcs ? c.hybridPhaseFunWidth : 0.0);
  f.put_att("Intensity_uses_Russian_roulette", NC_INT, (int)c.useRussianRouletteForIntensity);
  f.put_att("Intensity_Russian_roulette_zeta_min", NC_FLOAT, c.useRussianRouletteForIntensity ? c.zetaMin : 0.0);
  f.put_att("limited_intensity_contributions", NC_INT, (int)c.limitIntensityContributions);
  f.put_att("max_intensity_contribution", NC_FLOAT, c.limitIntensityContributions ? c.maxIntensityContribution : 0.0);
  f.put_att("Cpu_time_total", NC_FLOAT, cpuTotal);
  f.put_att("Cpu_time_setup", NC_FLOAT, cpuSetup);
  f.put_att("Number_of_processors_used", NC_INT, numProcs);
  const bool prof = c.reportAbsorptionProfile, vol = c.reportVolumeAbsorption && !s.absorbedVolume.empty();
  int xd = f.def_dim("x", (uint32_t)nx), yd = f.def_dim("y", (uint32_t)ny), zd = -1;
  if (prof || vol) zd = f.def_dim("z", (uint32_t)nz);
  auto mids = [](const std::vector<float>& e) {
    std::vector<float> m(e.size() - 1);
    for (size_t i = 0; i + 1 < e.size(); i++) m[i] = (e[i] + e[i + 1]) / 2;
    return m;
  };
  auto xm = mids(x), ym = mids(y), zm = mids(z);
  f.def_var("x", NC_FLOAT, {xd}, xm.data());
  f.def_var("y", NC_FLOAT, {yd}, ym.data());
  if (prof || vol) f.def_var("z", NC_FLOAT, {zd}, zm.data());
  f.def_var("fluxUp", NC_FLOAT, {yd, xd}, s.fluxUp.data());
  f.def_var("fluxDown", NC_FLOAT, {yd, xd}, s.fluxDown.data());
  f.def_var("fluxAbsorbed", NC_FLOAT, {yd, xd}, s.fluxAbsorbed.data());
  f.def_var("fluxUp_StdErr", NC_FLOAT, {yd, xd}, s.fluxUp.data() + ncol);
  f.def_var("fluxDown_StdErr", NC_FLOAT, {yd, xd}, s.fluxDown.data() + ncol);
  f.def_var("fluxAbsorbed_StdErr", NC_FLOAT, {yd, xd}, s.fluxAbsorbed.data() + ncol);
  if (prof) {
    f.def_var("absorptionProfile", NC_FLOAT, {zd}, s.absorbedProfile.data());
    f.def_var("absorptionProfile_StdErr", NC_FLOAT, {zd}, s.absorbedProfile.data() + nz);
  }
  if (vol) {
    f.def_var("absorbedVolume", NC_FLOAT, {zd, yd, xd}, s.absorbedVolume.data());
    f.def_var("absorbedVolume_StdErr", NC_FLOAT, {zd, yd, xd}, s.absorbedVolume.data() + ncell);
  }
  if (computeIntensity && !s.radiance.empty()) {
    const int nd = (int)c.intensityMus.size();
    int dd = f.def_dim("direction", (uint32_t)nd);
    f.def_var("intensityMus", NC_FLOAT, {dd}, c.intensityMus.data());
    f.def_var("intensityPhis", NC_FLOAT, {dd}, c.intensityPhis.data());
    f.def_var("intensity", NC_FLOAT, {dd, yd, xd}, s.radiance.data());
    f.def_var("intensity_StdErr", NC_FLOAT, {dd, yd, xd}, s.radiance.data() + (size_t)nd * ncol);
  }
  return f.write(c.outputNetcdfFile);
}

}  // namespace i3rc_host
