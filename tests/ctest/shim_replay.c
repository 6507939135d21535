/* Replays, in plain C, exactly the calls the Fortran ISO_C_BINDING shim (the .f95 files under fortran/) makes on the C ABI of
 * include/i3rc_b200.h when a reference driver runs through it (Example-Drivers/monteCarloDriver.f95:169-326):
 *
 *   new_Integrator(domain)            -> i3rc_new_Integrator (dense arrays of getOpticalPropertiesByComponent) +
 *                                        i3rc_set_phase_table per component             (shim: new_Integrator)
 *   specifyParameters x 3             -> i3rc_specifyParameters with the presence masks of monteCarloDriver.f95:185-216
 *   per batch: new_RandomNumberSequence((/ iseed, batch /)), new_PhotonStream(solarMu, solarAzimuth, N, randoms)
 *                                     -> a descriptor, no host draws                     (shim: monteCarloIllumination)
 *              computeRadiativeTransfer -> i3rc_computeRadiativeTransfer(handle, &descriptor, seed, 2)
 *              reportResults            -> i3rc_reportResults(NULL = absent optional)
 *   finalize_Integrator               -> i3rc_finalize_Integrator
 *
 * The problem: a 4 x 2 x 3 slab, Henyey-Greenstein g = 0.85 by 16 Legendre coefficients, ssa 0.95, one radiance direction.
 * Prints one line per batch (meanFluxUp meanFluxDown meanFluxAbsorbed meanIntensity) for the test-suite to compare with
 * the Python mirror's run of the same problem.  Test infrastructure; no Fortran compiler exists in this image.
 * Exit codes: 0 ok, 3 no CUDA device (the message of i3rc_last_message is printed), 1 anything else. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "i3rc_b200.h"

#define NX 4
#define NY 2
#define NZ 3
#define NCOEF 16

static int fail(const char* what, i3rc_integrator* h) {
  fprintf(stderr, "%s: %s\n", what, i3rc_last_message(h));
  return 1;
}

int main(int argc, char** argv) {
  const int nBatches = argc > 1 ? atoi(argv[1]) : 4, nPhotons = argc > 2 ? atoi(argv[2]) : 20000, iseed = 10;
  float x[NX + 1], y[NY + 1], z[NZ + 1];
  for (int i = 0; i <= NX; i++) x[i] = 125.0f * i;
  for (int i = 0; i <= NY; i++) y[i] = 250.0f * i;
  for (int i = 0; i <= NZ; i++) z[i] = 100.0f * i;
  /* getOpticalPropertiesByComponent (Code/opticalProperties.f95:429-539), one component: Fortran order, x fastest */
  static float totalExt[NX * NY * NZ], cumExt[NX * NY * NZ], ssa[NX * NY * NZ];
  static int32_t pf[NX * NY * NZ];
  for (int k = 0; k < NZ; k++)
    for (int j = 0; j < NY; j++)
      for (int i = 0; i < NX; i++) {
        const int c = (k * NY + j) * NX + i;
        totalExt[c] = (i < 2 ? 0.004f : 0.012f) * (1.0f + 0.25f * k);
        cumExt[c] = 1.0f;
        ssa[c] = 0.95f;
        pf[c] = 1;
      }
  i3rc_integrator* h = NULL;
  if (i3rc_new_Integrator(NX, NY, NZ, 1, x, y, z, totalExt, cumExt, ssa, pf, &h) == I3RC_FAILURE) {
    fprintf(stderr, "new_Integrator: %s\n", i3rc_last_message(NULL));
    return i3rc_device_count() == 0 ? 3 : 1;
  }
  /* passPhaseTable: Legendre coefficients of the one table entry */
  float coefs[NCOEF];
  int32_t offsets[2] = {0, NCOEF};
  for (int l = 0; l < NCOEF; l++) coefs[l] = powf(0.85f, (float)(l + 1));
  i3rc_phase_table t = {1, 1, offsets, coefs, 0, NULL, NULL};
  if (i3rc_set_phase_table(h, 0, &t) == I3RC_FAILURE) return fail("set_phase_table", h);

  /* monteCarloDriver.f95:185-216: three specifyParameters calls */
  i3rc_params p = {0};
  p.present = I3RC_P_surfaceAlbedo | I3RC_P_minInverseTableSize;
  p.surfaceAlbedo = 0.2f;
  p.minInverseTableSize = 10001;
  if (i3rc_specifyParameters(h, &p) == I3RC_FAILURE) return fail("specifyParameters", h);
  float mus[1] = {1.0f}, phis[1] = {0.0f};
  i3rc_params q = {0};
  q.present = I3RC_P_minForwardTableSize | I3RC_P_intensityMus | I3RC_P_intensityPhis | I3RC_P_computeIntensity;
  q.minForwardTableSize = 10001;
  q.numIntensityDirections = 1;
  q.intensityMus = mus;
  q.intensityPhis = phis;
  q.computeIntensity = 1;
  if (i3rc_specifyParameters(h, &q) == I3RC_FAILURE) return fail("specifyParameters", h);
  i3rc_params r = {0};
  r.present = I3RC_P_useRayTracing | I3RC_P_useRussianRoulette | I3RC_P_useRussianRouletteForIntensity | I3RC_P_zetaMin |
              I3RC_P_useHybridPhaseFunsForIntenCalcs | I3RC_P_hybridPhaseFunWidth | I3RC_P_numOrdersOrigPhaseFunIntenCalcs |
              I3RC_P_limitIntensityContributions | I3RC_P_maxIntensityContribution;
  r.useRayTracing = 1;
  r.useRussianRoulette = 1;
  r.useRussianRouletteForIntensity = 1;
  r.zetaMin = 0.3f;
  r.hybridPhaseFunWidth = 0.0f;
  r.maxIntensityContribution = 0.0f;
  if (i3rc_specifyParameters(h, &r) == I3RC_FAILURE) return fail("specifyParameters", h);
  if (!i3rc_isReady_Integrator(h)) return fail("isReady_Integrator", h);

  for (int batch = 1; batch <= nBatches; batch++) {
    int32_t seed[2] = {iseed, batch};          /* randomNumbers%seed of the replacement module RandomNumbers */
    i3rc_photon_source src = {0};              /* photons%descriptor of newPhotonStream_Directional */
    src.kind = I3RC_SRC_DIRECTIONAL;
    src.numberOfPhotons = nPhotons;
    src.solarMu = 0.5f;
    src.solarAzimuth = 0.0f;
    if (i3rc_computeRadiativeTransfer(h, &src, seed, 2) == I3RC_FAILURE) return fail("computeRadiativeTransfer", h);
    float up, down, absorbed, meanI[1], fluxUp[NX * NY], inten[NX * NY];
    if (i3rc_reportResults(h, &up, &down, &absorbed, fluxUp, NULL, NULL, NULL, NULL, meanI, inten) == I3RC_FAILURE)
      return fail("reportResults", h);
    float colMean = 0.0f;
    for (int c = 0; c < NX * NY; c++) colMean += fluxUp[c] / (NX * NY);
    printf("%d %.7f %.7f %.7f %.7f %.7f\n", batch, up, down, absorbed, meanI[0], colMean);
  }
  i3rc_finalize_Integrator(h);
  return 0;
}
