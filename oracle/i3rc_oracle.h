/*
 * oracle/i3rc_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the reference (Fortran 95) cannot be compiled in this environment (no Fortran
 * compiler, no netCDF-Fortran, no MPI) and ships no golden vectors of any kind.  The oracle is a
 * restatement of the reference's algorithm; the only external pin is the MT19937 stream (public
 * known-answer vectors) plus analytic known answers (see tests/test_oracle_pins.py).
 *
 * The structs below are laid out identically to the ones in include/i3rc_b200.h so that one set
 * of ctypes definitions drives both libraries.
 */
#ifndef I3RC_ORACLE_H
#define I3RC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_integrator orc_integrator;

typedef struct {
  int32_t kind;                /* 1 = Legendre coefficients, 2 = tabulated on one angle set */
  int32_t n_entries;
  const int32_t* coef_offsets; /* [n_entries+1] into coefs (Legendre) */
  const float* coefs;          /* coefficients starting with P1 (P0 == 1 implied) */
  int32_t n_angles;            /* tabulated */
  const float* angles;         /* [n_angles] radians, 0..pi */
  const float* values;         /* [n_entries][n_angles], angle fastest */
} orc_phase_table;

typedef struct {
  const float* extinction;     /* (nxc,nyc,nz) x fastest; nxc,nyc = nx,ny or 1,1 */
  const float* ssa;
  const int32_t* phase_index;  /* 1-based entry in table, 0 where no extinction */
  int32_t horizontally_uniform;
  int32_t z_level_base;        /* 1-based */
  int32_t nz;
  orc_phase_table table;
} orc_component;

typedef struct {
  uint32_t present;            /* bit mask of I3RC_P_* */
  float surfaceAlbedo;
  int32_t minForwardTableSize, minInverseTableSize;
  int32_t numIntensityDirections;
  const float* intensityMus;
  const float* intensityPhis;  /* degrees */
  int32_t computeIntensity, useRayTracing, useRussianRoulette, useRussianRouletteForIntensity;
  float zetaMin;
  int32_t useHybridPhaseFunsForIntenCalcs;
  float hybridPhaseFunWidth;
  int32_t numOrdersOrigPhaseFunIntenCalcs;
  int32_t limitIntensityContributions;
  float maxIntensityContribution;
  int32_t surf_nx, surf_ny;    /* surfaceBDRF (Lambertian map): cells */
  const float* surf_x;         /* [surf_nx+1] */
  const float* surf_y;         /* [surf_ny+1] */
  const float* surf_params;    /* [surf_ny][surf_nx] albedo */
} orc_params;

enum {
  I3RC_P_surfaceAlbedo = 1u << 0,
  I3RC_P_surfaceBDRF = 1u << 1,
  I3RC_P_minForwardTableSize = 1u << 2,
  I3RC_P_minInverseTableSize = 1u << 3,
  I3RC_P_intensityMus = 1u << 4,
  I3RC_P_intensityPhis = 1u << 5,
  I3RC_P_computeIntensity = 1u << 6,
  I3RC_P_useRayTracing = 1u << 7,
  I3RC_P_useRussianRoulette = 1u << 8,
  I3RC_P_useRussianRouletteForIntensity = 1u << 9,
  I3RC_P_zetaMin = 1u << 10,
  I3RC_P_useHybridPhaseFunsForIntenCalcs = 1u << 11,
  I3RC_P_hybridPhaseFunWidth = 1u << 12,
  I3RC_P_numOrdersOrigPhaseFunIntenCalcs = 1u << 13,
  I3RC_P_limitIntensityContributions = 1u << 14,
  I3RC_P_maxIntensityContribution = 1u << 15
};

enum {
  I3RC_SRC_DIRECTIONAL = 1,
  I3RC_SRC_RANDOM_AZIMUTH = 2,
  I3RC_SRC_FLUX = 3,
  I3RC_SRC_SPOTLIGHT = 4,
  I3RC_SRC_INTERNAL_FLUX = 5,
  I3RC_SRC_INTERNAL_INTENSITY = 6,
  I3RC_SRC_ARRAYS = 7
};

typedef struct {
  int32_t kind;
  int32_t reserved;
  int64_t numberOfPhotons;
  float solarMu, solarAzimuth; /* azimuth in degrees */
  float x, y, z;               /* spotlight solarX/solarY; detector X/Y/Z (fractions of the domain) */
  float detectorMu, detectorPhi;
  int32_t detectorPointsUp;
  int32_t has_deltaX, has_deltaY;
  float deltaX, deltaY;
  const float* xPosition;      /* I3RC_SRC_ARRAYS: the five public arrays of type(photonStream) */
  const float* yPosition;
  const float* zPosition;
  const float* initialMu;
  const float* initialPhi;
} orc_photon_source;

typedef struct {
  int64_t photons, bad;
  int64_t crossings_photon, crossings_intensity;
  int64_t collisions, absorptions, contributions, exits_top, surface_hits;
  int64_t rng_draws, roulette_kills, null_collisions;
  /* cells a ray passed without gathering them (uniform slabs crossed in one go, empty-space codes), by photon path
   * segments and by local-estimate rays; 0 in the reference's algorithm: crossings_photon + cells_skipped is its number of
   * cell crossings of photon paths */
  int64_t cells_skipped, cells_skipped_intensity;
} orc_counters;

/* status codes shared with the product */
enum { I3RC_SUCCESS = 0, I3RC_WARNING = 1, I3RC_FAILURE = 2 };

/* --- RNG (RandomNumbersForMC.f95) --- */
void orc_mt_seed_vector(uint32_t* state625, const int32_t* seed, int n);
void orc_mt_seed_scalar(uint32_t* state625, int32_t seed);
uint32_t orc_mt_int32(uint32_t* state625);
float orc_mt_real(uint32_t* state625);

/* --- numericUtilities.f95 --- */
int orc_findIndex(float value, const float* table, int n, int firstGuess /* <=0: absent */);
void orc_computeLobattoMus(float* mus, int n);
void orc_computeLegendrePolynomials(int maxL, const float* mus, int nmu, float* P /* [nmu][maxL+1] */);

/* --- scatteringPhaseFunctions.f95 / inversePhaseFunctions.f95 --- */
int orc_phase_values_one(const orc_phase_table* t, int entry /*0-based*/, const float* angles, int n, float* out);
int orc_phase_values_table(const orc_phase_table* t, const float* angles, int n, float* out /*[n_entries][n]*/);
void orc_normalize_phase_function(const float* angles, const float* values, int n, float* out);
int orc_inverse_phase_function(const orc_phase_table* t, int entry, int nSteps, float* out);
void orc_hybrid_phase_functions(const float* angles, int nAngles, int nEntries, const float* values, float widthDeg, float* out);

/* --- opticalProperties.f95:429 --- */
int orc_getOpticalPropertiesByComponent(int nx, int ny, int nz, int nc, const orc_component* comps,
                                        float* totalExt, float* cumExt, float* ssa, int32_t* pfIndex);

/* --- integrator (monteCarloRadiativeTransfer.f95) --- */
int orc_new_Integrator(int nx, int ny, int nz, int nc, const float* xPos, const float* yPos, const float* zPos,
                       const float* totalExt, const float* cumExt, const float* ssa, const int32_t* pfIndex,
                       orc_integrator** out);
int orc_new_Integrator_components(int nx, int ny, int nz, const float* xPos, const float* yPos, const float* zPos,
                                  int nc, const orc_component* comps, orc_integrator** out);
int orc_set_phase_table(orc_integrator* h, int comp /*0-based*/, const orc_phase_table* t);
int orc_copy_Integrator(const orc_integrator* src, orc_integrator** out);
void orc_finalize_Integrator(orc_integrator* h);
int orc_isReady_Integrator(const orc_integrator* h);
int orc_specifyParameters(orc_integrator* h, const orc_params* p);
int orc_computeRadiativeTransfer(orc_integrator* h, const orc_photon_source* src, const int32_t* seed, int nseed);
int orc_reportResults(orc_integrator* h, float* meanFluxUp, float* meanFluxDown, float* meanFluxAbsorbed,
                      float* fluxUp, float* fluxDown, float* fluxAbsorbed, float* absorbedProfile,
                      float* volumeAbsorption, float* meanIntensity, float* intensity);
int orc_get_intensityByComponent(orc_integrator* h, float* out);
const char* orc_last_message(const orc_integrator* h);
void orc_get_counters(const orc_integrator* h, orc_counters* c);
int orc_get_table(orc_integrator* h, int which /*0 inverse,1 forward,2 forward-orig*/, int comp, float* out, int* nSteps, int* nEntries);
int orc_tabulate(orc_integrator* h);

/* deterministic sub-path probes */
int orc_trace_rays(orc_integrator* h, int n, const float* pos /*[n][3]*/, const float* dir /*[n][3]*/,
                   const float* tauLimit /* may be NULL */, float* tauOut, float* posOut, int32_t* idxOut);
int orc_sample_scattering_angles(orc_integrator* h, int comp, int entry, int n, const float* xi, float* theta);
int orc_lookup_phase_function(orc_integrator* h, int comp, int entry, int which, int n, const float* angles, float* out);
void orc_next_direct(const float* xi /*>= 2 per try*/, int nxi, float scatteringCosine, float* S);

/* batch driver (monteCarloDriver.f95:264-378); seedOrder 0: (iseed,batch)  1: (batch,iseed) */
typedef struct {
  int32_t nx, ny, nz, nd, with_volume;
  double* meanFluxUp;    /* [2] */
  double* meanFluxDown;
  double* meanFluxAbsorbed;
  double* fluxUp;        /* [2][ny][nx] */
  double* fluxDown;
  double* fluxAbsorbed;
  double* absorbedProfile; /* [2][nz] */
  double* absorbedVolume;  /* [2][nz][ny][nx] or NULL */
  double* radiance;        /* [2][nd][ny][nx] or NULL */
  double* meanRadiance;    /* [2][nd] or NULL */
} orc_batch_stats;
int orc_run_batches(const orc_integrator* proto, const orc_photon_source* src, int32_t iseed, int seedOrder,
                    int batchBegin, int nBatches, int nThreads, orc_batch_stats* stats, orc_counters* counters);

#ifdef __cplusplus
}
#endif
#endif
