/*
 * include/i3rc_b200.h -- C ABI of the B200-native I3RC Monte Carlo photon-tracing integrator.
 *
 * This is the drop-in boundary for the one hot path of the reference:
 *   Integrators/monteCarloRadiativeTransfer.f95 :: computeRadiativeTransfer  (MCRT below)
 * Every entry point names the reference procedure it replaces (file:line, relative to the reference
 * tree).  A Fortran module with the reference's name and public list (MCRT:154-156) binds these with
 * ISO_C_BINDING; the stub is shown in INTEGRATION.md.  All pointers are HOST pointers unless a name
 * says "dev"; arrays use the reference's Fortran order (x fastest, then y, z, component / direction).
 *
 * Return codes follow Code/ErrorMessages.f95: 0 success, 1 warning, 2 failure; the text that the
 * reference would push on its ErrorMessage stack is returned by i3rc_last_message().
 *
 * There is no CPU fallback: every compute entry point fails (I3RC_FAILURE, message "no CUDA device")
 * when no sm_100 GPU is present.
 */
#ifndef I3RC_B200_H
#define I3RC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { I3RC_SUCCESS = 0, I3RC_WARNING = 1, I3RC_FAILURE = 2 };

typedef struct i3rc_integrator i3rc_integrator; /* opaque: type(integrator), MCRT:50-142 */

/* One phaseFunctionTable (Code/scatteringPhaseFunctions.f95:48-58), either all-Legendre or tabulated on
 * one shared angle set (the two forms the reference can persist, scatteringPhaseFunctions.f95:905-907). */
typedef struct {
  int32_t kind;                /* 1 = Legendre coefficients, 2 = angle/value pairs on one angle set */
  int32_t n_entries;
  const int32_t* coef_offsets; /* [n_entries+1] offsets into coefs */
  const float* coefs;          /* Legendre coefficients starting at P1 (P0 == 1 is implied) */
  int32_t n_angles;
  const float* angles;         /* [n_angles] radians, first 0, last pi */
  const float* values;         /* [n_entries][n_angles], angle fastest; normalised on ingestion */
} i3rc_phase_table;

/* One opticalComponent (Code/opticalProperties.f95:34-52). */
typedef struct {
  const float* extinction;     /* (nx,ny,nz_c) or (1,1,nz_c) when horizontally_uniform */
  const float* ssa;
  const int32_t* phase_index;  /* 1-based entry of table; 0 where there is no extinction */
  int32_t horizontally_uniform;
  int32_t z_level_base;        /* 1-based first level, as zLevelBase */
  int32_t nz;                  /* nz_c */
  i3rc_phase_table table;
} i3rc_component;

/* Optional arguments of specifyParameters (MCRT:830-838); `present` says which are supplied. */
typedef struct {
  uint32_t present;            /* OR of I3RC_P_* */
  float surfaceAlbedo;
  int32_t minForwardTableSize, minInverseTableSize;
  int32_t numIntensityDirections;
  const float* intensityMus;
  const float* intensityPhis;  /* degrees */
  int32_t computeIntensity, useRayTracing, useRussianRoulette, useRussianRouletteForIntensity;
  float zetaMin;
  int32_t useHybridPhaseFunsForIntenCalcs;
  float hybridPhaseFunWidth;   /* degrees */
  int32_t numOrdersOrigPhaseFunIntenCalcs;
  int32_t limitIntensityContributions;
  float maxIntensityContribution;
  /* surfaceBDRF: a surfaceDescription (Code/surfaceProperties.f95:34-38), Lambertian albedo map */
  int32_t surf_nx, surf_ny;
  const float* surf_x;         /* [surf_nx+1] */
  const float* surf_y;         /* [surf_ny+1] */
  const float* surf_params;    /* [surf_ny][surf_nx] */
} i3rc_params;

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

/* A photonStream (Code/monteCarloIllumination.f95:34-50) as a DESCRIPTOR: the six constructors behind
 * new_PhotonStream become `kind`; photons are generated on the device from the Philox stream of the
 * batch.  I3RC_SRC_ARRAYS carries the five public host arrays of the reference type for drivers that
 * fill them by hand. */
enum {
  I3RC_SRC_DIRECTIONAL = 1,        /* monteCarloIllumination.f95:62  */
  I3RC_SRC_RANDOM_AZIMUTH = 2,     /* :106 */
  I3RC_SRC_FLUX = 3,               /* :148 */
  I3RC_SRC_SPOTLIGHT = 4,          /* :187 */
  I3RC_SRC_INTERNAL_FLUX = 5,      /* :228 */
  I3RC_SRC_INTERNAL_INTENSITY = 6, /* :329 */
  I3RC_SRC_ARRAYS = 7              /* :34-41 public components */
};
typedef struct {
  int32_t kind;
  int32_t reserved;
  int64_t numberOfPhotons;
  float solarMu, solarAzimuth;     /* azimuth in degrees */
  float x, y, z;                   /* spotlight solarX/solarY; detector X/Y/Z, as fractions of the domain */
  float detectorMu, detectorPhi;
  int32_t detectorPointsUp;
  int32_t has_deltaX, has_deltaY;
  float deltaX, deltaY;
  const float* xPosition;          /* I3RC_SRC_ARRAYS only (host, pinned or pageable) */
  const float* yPosition;
  const float* zPosition;
  const float* initialMu;
  const float* initialPhi;
} i3rc_photon_source;

/* Device event counters of the last computeRadiativeTransfer (accumulated until reset by the next call). */
typedef struct {
  int64_t photons, bad;
  int64_t crossings_photon, crossings_intensity;
  int64_t collisions, absorptions, contributions, exits_top, surface_hits;
  int64_t rng_draws, roulette_kills, null_collisions;
  /* cells a ray passed without gathering them (uniform slabs crossed in one go, empty-space codes), by photon path
   * segments and by local-estimate rays; 0 in the reference's algorithm: crossings_photon + cells_skipped is its number of
   * cell crossings of photon paths */
  int64_t cells_skipped, cells_skipped_intensity;
} i3rc_counters;

/* ---- device selection -------------------------------------------------------------------- */
int i3rc_device_count(void);
int i3rc_set_device(int device);                 /* one process per GPU: call before new_Integrator */

/* ---- integrator lifecycle ---------------------------------------------------------------- */
/* new_Integrator (MCRT:162-254) from the dense arrays getOpticalPropertiesByComponent returns
 * (Code/opticalProperties.f95:429-539): totalExt(nx,ny,nz), cumExt/ssa/pfIndex(nx,ny,nz,nc). */
int i3rc_new_Integrator(int nx, int ny, int nz, int nc, const float* xPos, const float* yPos, const float* zPos,
                        const float* totalExt, const float* cumExt, const float* ssa, const int32_t* pfIndex,
                        i3rc_integrator** out);
/* new_Integrator(domain): getOpticalPropertiesByComponent (opticalProperties.f95:429-539) runs on the device. */
int i3rc_new_Integrator_components(int nx, int ny, int nz, const float* xPos, const float* yPos, const float* zPos,
                                   int nc, const i3rc_component* comps, i3rc_integrator** out);
/* Spectral loops (the reference's own kDistribution.f95 is an unfinished stub): replace the extinction
 * of component `comp` (0-based) by a horizontally uniform profile extinction[nz] -- one k-distribution term of a gas
 * component -- without rebuilding or re-uploading the other fields.  Single-scattering albedo and phase-function index of
 * the component stay as they are.  Equivalent to replaceOpticalComponent (Code/opticalProperties.f95:232-309) followed
 * by new_Integrator. */
int i3rc_set_component_profile(i3rc_integrator* h, int comp, const float* extinction);
/* forwardTables(comp) (MCRT:92-93): the table tabulate{Inverse,Forward}PhaseFunctions work from. comp is 0-based. */
int i3rc_set_phase_table(i3rc_integrator* h, int comp, const i3rc_phase_table* t);
int i3rc_copy_Integrator(const i3rc_integrator* src, i3rc_integrator** out); /* MCRT:1082 */
void i3rc_finalize_Integrator(i3rc_integrator* h);                            /* MCRT:1258 */
int i3rc_isReady_Integrator(const i3rc_integrator* h);                        /* MCRT:1073 */
const char* i3rc_last_message(const i3rc_integrator* h);

/* ---- specifyParameters (MCRT:830-1069) ---------------------------------------------------- */
int i3rc_specifyParameters(i3rc_integrator* h, const i3rc_params* p);

/* ---- computeRadiativeTransfer (MCRT:262-398) ----------------------------------------------
 * `seed` is the vector the driver hands to new_RandomNumberSequence (monteCarloDriver.f95:277:
 * (/ iseed, batch /)); it keys the per-photon Philox4x32-10 streams that replace RandomNumbersForMC. */
int i3rc_computeRadiativeTransfer(i3rc_integrator* h, const i3rc_photon_source* src, const int32_t* seed, int nseed);

/* ---- reportResults (MCRT:711-826); NULL = optional argument absent ------------------------- */
int i3rc_reportResults(i3rc_integrator* h, float* meanFluxUp, float* meanFluxDown, float* meanFluxAbsorbed,
                       float* fluxUp, float* fluxDown, float* fluxAbsorbed, float* absorbedProfile,
                       float* volumeAbsorption, float* meanIntensity, float* intensity);
int i3rc_get_intensityByComponent(i3rc_integrator* h, float* out); /* (nx,ny,nDir,0:nc), MCRT:139-140 */
void i3rc_get_counters(const i3rc_integrator* h, i3rc_counters* c);

/* ---- phase-function tables (MCRT:1809-1998, Code/inversePhaseFunctions.f95:28-176) ---------- */
int i3rc_tabulate(i3rc_integrator* h); /* tabulateInverse/ForwardPhaseFunctions now (normally lazy) */
/* which: 0 inverse, 1 forward (hybrid if enabled), 2 forward original.  out may be NULL to query sizes. */
int i3rc_get_table(i3rc_integrator* h, int which, int comp, float* out, int* nSteps, int* nEntries);
/* inject externally computed tables (e.g. from the Fortran computeInversePhaseFuncTable) */
int i3rc_set_inverse_table(i3rc_integrator* h, int comp, int nSteps, int nEntries, const float* values);
int i3rc_set_forward_table(i3rc_integrator* h, int comp, int nSteps, int nEntries, const float* tabulated,
                           const float* original);

/* ---- deterministic sub-paths, exposed for parity tests ------------------------------------- */
/* accumulateExtinctionAlongPath (MCRT:1654-1807) for n rays; idxOut is 1-based like the reference. */
int i3rc_trace_rays(i3rc_integrator* h, int n, const float* pos, const float* dir, const float* tauLimit,
                    float* tauOut, float* posOut, int32_t* idxOut);
int i3rc_sample_scattering_angles(i3rc_integrator* h, int comp, int entry, int n, const float* xi,
                                  float* theta); /* computeScatteringAngle, MCRT:1390 */
int i3rc_lookup_phase_function(i3rc_integrator* h, int comp, int entry, int which, int n, const float* angles,
                               float* out);      /* lookUpPhaseFuncValsFromTable, MCRT:1613 */

/* The device's random-number stream (Philox4x32-10, one stream per photon: replaces Code/RandomNumbersForMC.f95:169-299)
 * for explicit (photon, block) counters: raw4[4n] words and u4[4n] deviates.  No handle needed. */
int i3rc_probe_philox(uint32_t key0, uint32_t key1, int n, const uint64_t* photon, const uint32_t* block,
                      uint32_t* raw4, float* u4);
/* next_direct (MCRT:2086-2113) as the transport kernel runs it, deviates from the photon's stream starting at `block`;
 * direction / newDirection are [n][3], blocksUsed[n] = Philox blocks consumed (two rejection rounds per block). */
int i3rc_probe_next_direct(uint32_t key0, uint32_t key1, int n, const uint64_t* photon, const uint32_t* block,
                           const float* direction, const float* cosine, float* newDirection, uint32_t* blocksUsed);

/* ---- roofline ceilings measured on the spot (no reference counterpart) ---------------- */
/* random 4-byte gathers per second over `bytes` of device memory (8 MB: L2-resident field; >= 1 GB: HBM) */
int i3rc_measure_gather_rate(size_t bytes, int repeats, double* gathersPerSec);
/* warp instructions per second the SMs issue when nothing stalls (independent FMAs) */
int i3rc_measure_issue_rate(int repeats, double* warpInstPerSec);

/* ---- batch moments on the device (monteCarloDriver.f95:300-378) ---------------------------- */
typedef struct {
  double* meanFluxUp;      /* [2]: mean, standard error */
  double* meanFluxDown;
  double* meanFluxAbsorbed;
  double* fluxUp;          /* [2][ny][nx] */
  double* fluxDown;
  double* fluxAbsorbed;
  double* absorbedProfile; /* [2][nz] */
  double* absorbedVolume;  /* [2][nz][ny][nx] or NULL */
  double* radiance;        /* [2][nDir][ny][nx] or NULL */
  double* meanRadiance;    /* [2][nDir] or NULL */
} i3rc_stats_out;
int i3rc_stats_reset(i3rc_integrator* h, int with_volume);
int i3rc_stats_accumulate(i3rc_integrator* h);  /* sum x and x**2 of every output of the last batch */
/* the packed double buffer of sums (device pointer) for the one collective that replaces the
 * sumAcrossProcesses calls (Code/multipleProcesses_mpi.f95:57-131) */
int i3rc_stats_device_buffer(i3rc_integrator* h, void** dev_ptr, int64_t* n_doubles);
/* mean = solarFlux*sum/nB ; stderr = sqrt(max(0, sum2/nB - mean^2)/(nB-1))  (monteCarloDriver.f95:358-378) */
int i3rc_stats_report(i3rc_integrator* h, double solarFlux, int numBatches, const i3rc_stats_out* out);
/* the batch loop of monteCarloDriver.f95:274-326 without host round trips: for b in
 * [batchBegin, batchBegin+nBatches): seed = seedOrder==0 ? (iseed,b) : (b,iseed); compute; accumulate. */
int i3rc_run_batches(i3rc_integrator* h, const i3rc_photon_source* src, int32_t iseed, int seedOrder,
                     int batchBegin, int nBatches);

/* ---- NCCL reduction of the batch moments (replaces multipleProcesses_mpi.f95) --------------- */
int i3rc_comm_unique_id(char* id128);  /* rank 0: ncclGetUniqueId, broadcast the 128 bytes yourself */
int i3rc_comm_init(i3rc_integrator* h, int nRanks, int rank, const char* id128);
int i3rc_stats_allreduce(i3rc_integrator* h); /* ncclAllReduce(sum, double) over the packed buffer */
int i3rc_comm_finalize(i3rc_integrator* h);

/* ---- streams, timing, tuning ----------------------------------------------------------------- */
int i3rc_synchronize(i3rc_integrator* h);
void* i3rc_stream(i3rc_integrator* h);          /* cudaStream_t all kernels of this handle run on */
/* device time (CUDA events on the handle's stream) of the transport kernels since the last reset */
/* layout decisions: what = 0: layers of totalExt stored in 3-D when horizontally uniform layers are kept out of the field
 * (0 = all), 1: floats of tallies staged per warp in shared memory (0 = global atomics only); -1 on error */
int i3rc_get_layout(i3rc_integrator* h, int what);
int i3rc_get_timing(i3rc_integrator* h, double* trace_ms, int64_t* trace_launches, int64_t* other_launches);
int i3rc_reset_timing(i3rc_integrator* h);
int i3rc_set_tuning(i3rc_integrator* h, const char* key, int value);
const char* i3rc_version(void);

#ifdef __cplusplus
}
#endif
#endif
