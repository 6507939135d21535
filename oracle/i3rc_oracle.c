/*
 * oracle/i3rc_oracle.c -- CPU restatement of the I3RC Monte Carlo photon-tracing integrator.
 *
 * TEST INFRASTRUCTURE ONLY (see i3rc_oracle.h).  PARITY UNPINNED: no Fortran compiler exists in
 * this environment and the reference ships no golden vectors; this file restates the reference's
 * algorithm in plain C, float32, in the reference's operation order, and replays the reference's
 * MT19937 stream bit for bit (pinned to the public mt19937ar known-answer vectors).
 *
 * Citations are to files under the reference tree:
 *   MCRT  = Integrators/monteCarloRadiativeTransfer.f95
 *   RNG   = Code/RandomNumbersForMC.f95
 *   NUM   = Code/numericUtilities.f95
 *   SPF   = Code/scatteringPhaseFunctions.f95
 *   IPF   = Code/inversePhaseFunctions.f95
 *   OPT   = Code/opticalProperties.f95
 *   ILL   = Code/monteCarloIllumination.f95
 *   SURF  = Code/surfaceProperties.f95
 *   DRV   = Example-Drivers/monteCarloDriver.f95
 *
 * Arrays keep the reference's 1-based indices (position vectors are allocated with a spare slot).
 * Compile with -ffp-contract=off so every float operation rounds like the default-kind REAL of the
 * reference (no -r8 anywhere in the reference's Makefile).
 */
#include "i3rc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Fortran intrinsics
 * ---------------------------------------------------------------------------------------- */
#define F_TINY FLT_MIN
#define F_HUGE FLT_MAX
static const float Pi = 3.14159265358979312f; /* MCRT:43; SPF:26 rounds to the same float */

static inline float f_spacing(float x) {
  if (x == 0.0f) return F_TINY;
  int e;
  (void)frexpf(fabsf(x), &e);
  float s = ldexpf(1.0f, e - 24);
  return s < F_TINY ? F_TINY : s;
}
static inline float f_sign(float a, float b) { return b >= 0.0f ? fabsf(a) : -fabsf(a); }

/* ------------------------------------------------------------------------------------------
 * RNG:25-299 -- MT19937 (mt19937ar), state[0..623] + cursor in state[624]
 * ---------------------------------------------------------------------------------------- */
#define MT_N 624
#define MT_M 397
void orc_mt_seed_scalar(uint32_t* mt, int32_t seed) { /* RNG:169-185 */
  mt[0] = (uint32_t)seed;
  for (int i = 1; i < MT_N; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  mt[MT_N] = MT_N;
}
void orc_mt_seed_vector(uint32_t* mt, const int32_t* seed, int n) { /* RNG:187-239 */
  orc_mt_seed_scalar(mt, 19650218);
  int i = 1, j = 0;
  int k = MT_N > n ? MT_N : n;
  for (; k; k--) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + (uint32_t)seed[j] + (uint32_t)j;
    i++;
    j++;
    if (i >= MT_N) {
      mt[0] = mt[MT_N - 1];
      i = 1;
    }
    if (j >= n) j = 0;
  }
  for (k = MT_N - 1; k; k--) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    i++;
    if (i >= MT_N) {
      mt[0] = mt[MT_N - 1];
      i = 1;
    }
  }
  mt[0] = 0x80000000u;
  mt[MT_N] = MT_N;
}
static void mt_next_state(uint32_t* mt) { /* RNG:134-152 */
#define MIXBITS(u, v) (((u)&0x80000000u) | ((v)&0x7fffffffu))
#define TWIST(u, v) ((MIXBITS(u, v) >> 1) ^ ((v)&1u ? 0x9908b0dfu : 0u))
  int k;
  for (k = 0; k < MT_N - MT_M; k++) mt[k] = mt[k + MT_M] ^ TWIST(mt[k], mt[k + 1]);
  for (; k < MT_N - 1; k++) mt[k] = mt[k + MT_M - MT_N] ^ TWIST(mt[k], mt[k + 1]);
  mt[MT_N - 1] = mt[MT_M - 1] ^ TWIST(mt[MT_N - 1], mt[0]);
  mt[MT_N] = 0;
}
uint32_t orc_mt_int32(uint32_t* mt) { /* RNG:243-258 */
  if (mt[MT_N] >= MT_N) mt_next_state(mt);
  uint32_t y = mt[mt[MT_N]++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
float orc_mt_real(uint32_t* mt) { /* RNG:275-299: genrand_real1 in double, cast to default real */
  double d = (double)orc_mt_int32(mt) / (4294967296.0 - 1.0);
  return (float)d;
}

/* ------------------------------------------------------------------------------------------
 * NUM:195-248 findIndex.  table is 1-based (table[1..n]); returns i with table(i) <= v < table(i+1),
 * 0 if v is below the table.
 * ---------------------------------------------------------------------------------------- */
static int findIndex1(float value, const float* table /*1-based*/, int n, int firstGuess) {
  int lowerBound, upperBound, midPoint, increment;
  if (firstGuess > 0) {
    lowerBound = firstGuess;
    increment = 1;
    for (;;) {
      upperBound = lowerBound + increment < n ? lowerBound + increment : n;
      if (lowerBound == n || (table[lowerBound] <= value && table[upperBound] > value)) break;
      if (table[lowerBound] > value) {
        upperBound = lowerBound;
        lowerBound = upperBound - increment > 1 ? upperBound - increment : 1;
      } else {
        lowerBound = upperBound;
      }
      increment *= 2;
    }
  } else {
    lowerBound = 0;
    upperBound = n;
  }
  for (;;) {
    if (lowerBound == n || upperBound <= lowerBound + 1) break;
    midPoint = (lowerBound + upperBound) / 2;
    if (value >= table[midPoint])
      lowerBound = midPoint;
    else
      upperBound = midPoint;
  }
  return lowerBound;
}
int orc_findIndex(float value, const float* table, int n, int firstGuess) {
  return findIndex1(value, table - 1, n, firstGuess);
}

/* NUM:175-193; P is [nmu][maxL+1] (l fastest) */
void orc_computeLegendrePolynomials(int maxL, const float* mus, int nmu, float* P) {
  for (int m = 0; m < nmu; m++) {
    float* p = P + (size_t)m * (maxL + 1);
    p[0] = 1.0f;
    if (maxL >= 1) p[1] = mus[m];
    for (int l = 1; l <= maxL - 1; l++)
      p[l + 1] = (((float)(2 * l + 1) * mus[m]) * p[l] - (float)l * p[l - 1]) / (float)(l + 1);
  }
}

/* NUM:15-102 computeLobattoTerms -- only the abscissas are used on the path (IPF:111-112) */
void orc_computeLobattoMus(float* mus /*0-based out*/, int nTerms) {
  const float relativeAccuracy = 3.0f;
  const int maxIterations = 25;
  float pi = acosf(-1.0f);
  int midPoint = (nTerms + 1) / 2;
  int nr = midPoint - 1;
  float* trialMus = (float*)malloc(sizeof(float) * (nr > 0 ? nr : 1));
  float* lastGuess = (float*)malloc(sizeof(float) * (nr > 0 ? nr : 1));
  float* legP = (float*)malloc(sizeof(float) * (size_t)(nr > 0 ? nr : 1) * nTerms);
  float c1 = (nTerms % 2 == 1) ? 1.0f : 0.5f;
  for (int i = 1; i <= nr; i++) trialMus[i - 1] = sinf(pi * ((float)i - c1) / ((float)nTerms - 1.0f + 0.5f));
  int maxL = nTerms - 1;
  /* first Newton iteration (NUM:47-57) */
  orc_computeLegendrePolynomials(maxL, trialMus, nr, legP);
  for (int i = 0; i < nr; i++) {
    const float* p = legP + (size_t)i * nTerms;
    float t = trialMus[i];
    float deriv = (float)(nTerms - 1) * (t * p[nTerms - 1] - p[nTerms - 2]) / (t * t - 1.0f);
    float second = (2.0f * t * deriv - ((float)(nTerms * (nTerms - 1)) * p[nTerms - 1])) / (1.0f - t * t);
    lastGuess[i] = t;
    trialMus[i] = t - deriv / second;
  }
  int it = 0;
  for (;;) {
    int allConverged = 1;
    for (int i = 0; i < nr; i++)
      if (!(fabsf(trialMus[i] - lastGuess[i]) <= relativeAccuracy * f_spacing(trialMus[i]))) allConverged = 0;
    if (allConverged) break;
    orc_computeLegendrePolynomials(maxL, trialMus, nr, legP);
    for (int i = 0; i < nr; i++) {
      if (fabsf(trialMus[i] - lastGuess[i]) > relativeAccuracy * f_spacing(trialMus[i])) {
        const float* p = legP + (size_t)i * nTerms;
        float t = trialMus[i];
        float deriv = (float)(nTerms - 1) * (t * p[nTerms - 1] - p[nTerms - 2]) / (t * t - 1.0f);
        float second = (2.0f * t * deriv - ((float)(nTerms * (nTerms - 1)) * p[nTerms - 1])) / (1.0f - t * t);
        lastGuess[i] = t;
        trialMus[i] = t - deriv / second;
      }
    }
    it++;
    if (it > maxIterations) break;
  }
  /* NUM:86-99 (1-based m[]) */
  float* m = mus - 1;
  m[1] = -1.0f;
  for (int i = 1; i <= nr; i++) m[midPoint - i + 1] = -trialMus[i - 1];
  if (nTerms % 2 == 0) {
    for (int i = 1; i <= midPoint; i++) m[midPoint + i] = -m[midPoint - i + 1];
  } else {
    for (int i = 0; i <= midPoint - 1; i++) m[midPoint + i] = -m[midPoint - i];
  }
  free(trialMus);
  free(lastGuess);
  free(legP);
}

/* ------------------------------------------------------------------------------------------
 * Phase functions (SPF)
 * ---------------------------------------------------------------------------------------- */
void orc_normalize_phase_function(const float* a, const float* v, int n, float* out) { /* SPF:1329-1345 */
  float dot = 0.0f;
  for (int i = 0; i < n - 1; i++) dot += (cosf(a[i + 1]) - cosf(a[i])) * (0.5f * (v[i + 1] + v[i]));
  for (int i = 0; i < n; i++) out[i] = -v[i] * 2.0f / dot;
}

/* SPF:446-529 getPhaseFunctionValues_one */
int orc_phase_values_one(const orc_phase_table* t, int entry, const float* angles, int n, float* out) {
  if (t->kind == 1) {
    int off = t->coef_offsets[entry];
    int maxL = t->coef_offsets[entry + 1] - off;
    if (maxL == 0) { /* SPF:484-489 (quirk Q5) */
      for (int i = 0; i < n; i++) out[i] = 1.0f / 2.0f;
      return 0;
    }
    float* P = (float*)malloc(sizeof(float) * (maxL + 1));
    for (int i = 0; i < n; i++) {
      float mu = cosf(angles[i]);
      orc_computeLegendrePolynomials(maxL, &mu, 1, P);
      /* matmul((/1, coeffs/) * (/(2l+1)/), legendreP) -- sequential in l (SPF:493-494) */
      float s = 0.0f;
      for (int l = 0; l <= maxL; l++) {
        float c = (l == 0) ? 1.0f : t->coefs[off + l - 1];
        s += (c * (float)(2 * l + 1)) * P[l];
      }
      out[i] = s;
    }
    free(P);
    return 0;
  }
  /* tabulated: interpolate linearly in cosine of the scattering angle (SPF:497-525) */
  int nStored = t->n_angles;
  const float* sa = t->angles - 1;                               /* 1-based */
  const float* val = t->values + (size_t)entry * nStored - 1;    /* 1-based */
  for (int i = 0; i < n; i++) {
    int idx = findIndex1(angles[i], sa, nStored, 0);
    int idx1 = idx + 1;
    float dMu;
    if (idx < nStored) {
      dMu = cosf(sa[idx1]) - cosf(sa[idx]);
    } else {
      dMu = F_HUGE;
      idx1 = idx;
    }
    float w = 1.0f - (cosf(angles[i]) - cosf(sa[idx])) / dMu;
    out[i] = w * val[idx] + (1.0f - w) * val[idx1];
  }
  return 0;
}

/* SPF:531-648 getPhaseFunctionValues_table; out is [n_entries][n] */
int orc_phase_values_table(const orc_phase_table* t, const float* angles, int n, float* out) {
  if (t->kind == 1) {
    int maxL = 0;
    for (int e = 0; e < t->n_entries; e++) {
      int L = t->coef_offsets[e + 1] - t->coef_offsets[e];
      if (L > maxL) maxL = L;
    }
    float* P = (float*)malloc(sizeof(float) * (maxL + 2));
    for (int i = 0; i < n; i++) {
      if (maxL > 0) {
        float mu = cosf(angles[i]);
        orc_computeLegendrePolynomials(maxL, &mu, 1, P);
        for (int l = 0; l <= maxL; l++) P[l] = (float)(2 * l + 1) * P[l]; /* SPF:582-583 */
      }
      for (int e = 0; e < t->n_entries; e++) {
        int off = t->coef_offsets[e];
        int L = t->coef_offsets[e + 1] - off;
        float s;
        if (L == 0) {
          s = 1.0f / 2.0f; /* SPF:622-623 */
        } else {
          s = 0.0f;
          for (int l = 0; l <= L; l++) s += ((l == 0) ? 1.0f : t->coefs[off + l - 1]) * P[l];
        }
        out[(size_t)e * n + i] = s;
      }
    }
    free(P);
    return 0;
  }
  int nStored = t->n_angles;
  const float* sa = t->angles - 1;
  int prev = 0;
  for (int i = 0; i < n; i++) {
    int idx = findIndex1(angles[i], sa, nStored, prev); /* SPF:591-597 */
    prev = idx;
    int idx1 = idx + 1;
    float dMu;
    if (idx < nStored) {
      dMu = cosf(sa[idx1]) - cosf(sa[idx]);
    } else {
      dMu = F_HUGE;
      idx1 = idx;
    }
    float w = 1.0f - (cosf(angles[i]) - cosf(sa[idx])) / dMu;
    for (int e = 0; e < t->n_entries; e++) {
      const float* val = t->values + (size_t)e * nStored - 1;
      out[(size_t)e * n + i] = w * val[idx] + (1.0f - w) * val[idx1];
    }
  }
  return 0;
}

/* IPF:68-176 computeInversePhaseFunction */
int orc_inverse_phase_function(const orc_phase_table* t, int entry, int nSteps, float* inverseTable0) {
  int nAngles;
  float *values, *mus, *cdf;
  float* inverseTable = inverseTable0 - 1; /* 1-based */
  if (t->kind == 2) {
    nAngles = t->n_angles;
    values = (float*)malloc(sizeof(float) * (nAngles + 1));
    mus = (float*)malloc(sizeof(float) * (nAngles + 1));
    cdf = (float*)malloc(sizeof(float) * (nAngles + 1));
    float* tmp = (float*)malloc(sizeof(float) * nAngles);
    orc_phase_values_one(t, entry, t->angles, nAngles, tmp); /* IPF:97-99 */
    for (int i = 1; i <= nAngles; i++) { /* IPF:100 */
      mus[i] = cosf(t->angles[nAngles - i]);
      values[i] = tmp[nAngles - i];
    }
    free(tmp);
  } else {
    int nMoments = t->coef_offsets[entry + 1] - t->coef_offsets[entry];
    nAngles = nMoments > 2 ? nMoments : 2; /* IPF:109 */
    values = (float*)malloc(sizeof(float) * (nAngles + 1));
    mus = (float*)malloc(sizeof(float) * (nAngles + 1));
    cdf = (float*)malloc(sizeof(float) * (nAngles + 1));
    orc_computeLobattoMus(mus + 1, nAngles); /* IPF:112 */
    float* ang = (float*)malloc(sizeof(float) * nAngles);
    float* tmp = (float*)malloc(sizeof(float) * nAngles);
    for (int i = 0; i < nAngles; i++) ang[i] = acosf(mus[nAngles - i]); /* IPF:113 */
    orc_phase_values_one(t, entry, ang, nAngles, tmp);
    for (int i = 1; i <= nAngles; i++) values[i] = tmp[nAngles - i]; /* IPF:114 */
    free(ang);
    free(tmp);
  }
  cdf[1] = 0.0f; /* IPF:122-129 */
  for (int i = 2; i <= nAngles; i++) cdf[i] = cdf[i - 1] + (mus[i] - mus[i - 1]) * 0.5f * (values[i] + values[i - 1]);
  float cdfN = cdf[nAngles];
  for (int i = 1; i <= nAngles; i++) cdf[i] = cdf[i] / cdfN;

  int* indicies = (int*)malloc(sizeof(int) * (nSteps + 1));
  indicies[1] = findIndex1(0.0f, cdf, nAngles, 0); /* IPF:133-137 */
  for (int i = 2; i <= nSteps; i++) {
    float p = (float)(i - 1) / (float)(nSteps - 1);
    indicies[i] = findIndex1(p, cdf, nAngles, indicies[i - 1]);
  }
  for (int i = 1; i <= nSteps - 1; i++) { /* IPF:139-169 */
    float p = (float)(i - 1) / (float)(nSteps - 1);
    int k = indicies[i];
    if (cdf[k + 1] - cdf[k] <= f_spacing(cdf[k])) {
      inverseTable[i] = acosf(mus[k]);
    } else if (fabsf(values[k] - values[k + 1]) <= f_spacing(values[k])) {
      inverseTable[i] = acosf(mus[k] + (mus[k + 1] - mus[k]) * (p - cdf[k]) / (cdf[k + 1] - cdf[k]));
    } else {
      inverseTable[i] =
          acosf(mus[k] + (mus[k + 1] - mus[k]) / (values[k] - values[k + 1]) *
                             (values[k] - sqrtf(((cdf[k + 1] - p) * (values[k] * values[k]) +
                                                 (p - cdf[k]) * (values[k + 1] * values[k + 1])) /
                                                (cdf[k + 1] - cdf[k]))));
    }
  }
  inverseTable[nSteps] = 0.0f;
  free(indicies);
  free(values);
  free(mus);
  free(cdf);
  return 0;
}

/* MCRT:2000-2039 */
static float computeNormalization(const float* ac, const float* v, const float* g, int nAngles, int ti) {
  /* all 1-based */
  float IntegralGaus = 0.0f, IntegralOrig = 0.0f;
  for (int k = 1; k <= ti - 1; k++) IntegralGaus += (0.5f * (g[k] + g[k + 1])) * (ac[k] - ac[k + 1]);
  for (int k = ti; k <= nAngles - 1; k++) IntegralOrig += (0.5f * (v[k] + v[k + 1])) * (ac[k] - ac[k + 1]);
  if (IntegralOrig >= 2.0f) return 1.0f / IntegralGaus;
  return (2.0f - IntegralOrig) / IntegralGaus;
}
static float phaseFuncDiff(const float* ac, const float* v, const float* g, int nAngles, int ti) {
  float P0 = computeNormalization(ac, v, g, nAngles, ti);
  return P0 * g[ti] - v[ti];
}
/* MCRT:1925-1998 computeHydridPhaseFunctions; values/out are [nEntries][nAngles] */
void orc_hybrid_phase_functions(const float* angles0, int nAngles, int nEntries, const float* values0, float width,
                                float* out0) {
  const float* angles = angles0 - 1;
  float* gaus = (float*)malloc(sizeof(float) * (nAngles + 1));
  float* ac = (float*)malloc(sizeof(float) * (nAngles + 1));
  for (int i = 1; i <= nAngles; i++) {
    ac[i] = cosf(angles[i]);
    float r = angles[i] / (width * Pi / 180.0f);
    gaus[i] = expf(-(r * r));
  }
  memcpy(out0, values0, sizeof(float) * (size_t)nAngles * nEntries);
  for (int e = 0; e < nEntries; e++) {
    const float* v = values0 + (size_t)e * nAngles - 1;
    float* nv = out0 + (size_t)e * nAngles - 1;
    int lowerBound = findIndex1(width * Pi / 180.0f, angles, nAngles, 0) + 1;
    if (lowerBound >= nAngles - 2) break; /* exit entryLoop */
    float lowDiff = phaseFuncDiff(ac, v, gaus, nAngles, lowerBound);
    int increment = 1, upperBound;
    float upDiff;
    int noRoot = 0;
    for (;;) {
      upperBound = lowerBound + increment < nAngles - 1 ? lowerBound + increment : nAngles - 1;
      upDiff = phaseFuncDiff(ac, v, gaus, nAngles, upperBound);
      if (lowerBound == nAngles - 1) {
        noRoot = 1;
        break;
      }
      if (lowDiff * upDiff < 0.0f) break;
      lowerBound = upperBound;
      lowDiff = upDiff;
      increment *= 2;
    }
    if (noRoot) continue; /* cycle entryLoop */
    for (;;) {
      if (upperBound <= lowerBound + 1) break;
      int midPoint = (lowerBound + upperBound) / 2;
      float midDiff = phaseFuncDiff(ac, v, gaus, nAngles, midPoint);
      if (midDiff * upDiff < 0.0f) {
        lowerBound = midPoint;
        lowDiff = midDiff;
      } else {
        upperBound = midPoint;
        upDiff = midDiff;
      }
    }
    int ti = lowerBound;
    float P0 = computeNormalization(ac, v, gaus, nAngles, ti);
    for (int k = 1; k <= ti; k++) nv[k] = P0 * gaus[k];
    for (int k = ti + 1; k <= nAngles; k++) nv[k] = v[k];
  }
  free(gaus);
  free(ac);
}

/* ------------------------------------------------------------------------------------------
 * OPT:429-539 getOpticalPropertiesByComponent
 * ---------------------------------------------------------------------------------------- */
int orc_getOpticalPropertiesByComponent(int nx, int ny, int nz, int nc, const orc_component* comps, float* totalExt,
                                        float* cumExt, float* ssa, int32_t* pfIndex) {
  size_t ncell = (size_t)nx * ny * nz;
  memset(totalExt, 0, sizeof(float) * ncell);
  memset(cumExt, 0, sizeof(float) * ncell * nc);
  memset(ssa, 0, sizeof(float) * ncell * nc);
  memset(pfIndex, 0, sizeof(int32_t) * ncell * nc);
  for (int c = 0; c < nc; c++) {
    const orc_component* k = &comps[c];
    int minZ = k->z_level_base;
    if (minZ < 1 || minZ + k->nz - 1 > nz) return I3RC_FAILURE;
    for (int kz = 0; kz < k->nz; kz++)
      for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
          size_t dst = (size_t)c * ncell + ((size_t)(minZ - 1 + kz) * ny + j) * nx + i;
          size_t src = k->horizontally_uniform ? (size_t)kz : ((size_t)kz * ny + j) * nx + i;
          cumExt[dst] = k->extinction[src];
          ssa[dst] = k->ssa[src];
          pfIndex[dst] = k->phase_index[src];
        }
  }
  for (int c = 1; c < nc; c++)
    for (size_t i = 0; i < ncell; i++) cumExt[c * ncell + i] = cumExt[c * ncell + i] + cumExt[(c - 1) * ncell + i];
  memcpy(totalExt, cumExt + (size_t)(nc - 1) * ncell, sizeof(float) * ncell);
  for (int c = 0; c < nc; c++)
    for (size_t i = 0; i < ncell; i++)
      if (totalExt[i] > F_TINY) cumExt[c * ncell + i] = cumExt[c * ncell + i] / totalExt[i];
  return I3RC_SUCCESS;
}

/* ------------------------------------------------------------------------------------------
 * The integrator object (MCRT:50-142)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int numX, numY; /* numX = table steps, numY = entries */
  float* values;  /* [numY][numX] */
} matrix;

typedef struct {
  int kind, n_entries, n_angles;
  int32_t* coef_offsets;
  float* coefs;
  float* angles;
  float* values;
} owned_table;

struct orc_integrator {
  int readyToCompute, computeIntensity;
  int minForwardTableSize, minInverseTableSize;
  int useRayTracing, useRussianRoulette;
  float RussianRouletteW;
  float surfaceAlbedo;
  int xyRegularlySpaced, zRegularlySpaced;
  float deltaX, deltaY, deltaZ, x0, y0, z0;
  int nx, ny, nz, nc;
  float *xPosition, *yPosition, *zPosition; /* 1-based: [1..n+1] */
  float* totalExt;
  float* cumulativeExt;
  float* ssa;
  int32_t* phaseFunctionIndex;
  int useSurfaceBDRF;
  int surf_nx, surf_ny;
  float *surf_x, *surf_y, *surf_params; /* surf_x/y 1-based */
  owned_table* forwardTables;           /* [nc] */
  matrix *tabulatedPhaseFunctions, *tabulatedOrigPhaseFunctions, *inversePhaseFunctions;
  int nDir;
  float* intensityDirections; /* [nDir][3] */
  int useHybridPhaseFunsForIntenCalcs;
  float hybridPhaseFunWidth;
  int numOrdersOrigPhaseFunIntenCalcs;
  int useRussianRouletteForIntensity;
  float zetaMin;
  int limitIntensityContributions;
  float maxIntensityContribution;
  float* intensityExcess; /* [(nc+1)][nDir] */
  float *fluxUp, *fluxDown, *fluxAbsorbed, *volumeAbsorption, *intensity, *intensityByComponent;
  orc_counters cnt;
  char message[512];
};

#define TE(I, ix, iy, iz) (I)->totalExt[((size_t)((iz)-1) * (I)->ny + ((iy)-1)) * (I)->nx + ((ix)-1)]
#define F4(I, arr, ix, iy, iz, c) \
  (I)->arr[(size_t)((c)-1) * (I)->nx * (I)->ny * (I)->nz + ((size_t)((iz)-1) * (I)->ny + ((iy)-1)) * (I)->nx + ((ix)-1)]
#define F2(I, arr, ix, iy) (I)->arr[(size_t)((iy)-1) * (I)->nx + ((ix)-1)]

static void set_msg(orc_integrator* I, const char* m) {
  strncpy(I->message, m, sizeof(I->message) - 1);
  I->message[sizeof(I->message) - 1] = 0;
}
const char* orc_last_message(const orc_integrator* h) { return h->message; }
void orc_get_counters(const orc_integrator* h, orc_counters* c) { *c = h->cnt; }

static float* dupf(const float* p, size_t n) {
  float* q = (float*)malloc(sizeof(float) * (n ? n : 1));
  if (p) memcpy(q, p, sizeof(float) * n);
  return q;
}
static float* edges1(const float* p, int n) { /* returns 1-based pointer */
  float* q = (float*)malloc(sizeof(float) * (n + 2));
  memcpy(q + 1, p, sizeof(float) * n);
  q[0] = 0.0f;
  q[n + 1] = 0.0f;
  return q;
}

int orc_new_Integrator(int nx, int ny, int nz, int nc, const float* xPos, const float* yPos, const float* zPos,
                       const float* totalExt, const float* cumExt, const float* ssa, const int32_t* pfIndex,
                       orc_integrator** out) { /* MCRT:162-254 */
  orc_integrator* I = (orc_integrator*)calloc(1, sizeof(orc_integrator));
  I->minForwardTableSize = 9001;
  I->minInverseTableSize = 9001;
  I->useRayTracing = 1;
  I->useRussianRoulette = 1;
  I->RussianRouletteW = 1.0f;
  I->hybridPhaseFunWidth = 7.0f;
  I->zetaMin = 0.3f;
  I->maxIntensityContribution = F_HUGE;
  I->nx = nx;
  I->ny = ny;
  I->nz = nz;
  I->nc = nc;
  I->xPosition = edges1(xPos, nx + 1);
  I->yPosition = edges1(yPos, ny + 1);
  I->zPosition = edges1(zPos, nz + 1);
  I->x0 = I->xPosition[1];
  I->y0 = I->yPosition[1];
  I->z0 = I->zPosition[1];
  float deltaX = I->xPosition[2] - I->xPosition[1];
  float deltaY = I->yPosition[2] - I->yPosition[1];
  float deltaZ = I->zPosition[2] - I->zPosition[1];
  int xyReg = 1, zReg = 1;
  for (int i = 1; i <= nx; i++)
    if (!(fabsf((I->xPosition[i + 1] - I->xPosition[i]) - deltaX) <= 2.0f * f_spacing(I->xPosition[i + 1]))) xyReg = 0;
  for (int i = 1; i <= ny; i++)
    if (!(fabsf((I->yPosition[i + 1] - I->yPosition[i]) - deltaY) <= 2.0f * f_spacing(I->yPosition[i + 1]))) xyReg = 0;
  for (int i = 1; i <= nz; i++)
    if (!(fabsf((I->zPosition[i + 1] - I->zPosition[i]) - deltaZ) <= f_spacing(I->zPosition[i + 1]))) zReg = 0;
  if (xyReg) {
    I->xyRegularlySpaced = 1;
    I->deltaX = deltaX;
    I->deltaY = deltaY;
  }
  if (zReg) {
    I->zRegularlySpaced = 1;
    I->deltaZ = deltaZ;
  }
  size_t ncell = (size_t)nx * ny * nz;
  I->totalExt = dupf(totalExt, ncell);
  I->cumulativeExt = dupf(cumExt, ncell * nc);
  I->ssa = dupf(ssa, ncell * nc);
  I->phaseFunctionIndex = (int32_t*)malloc(sizeof(int32_t) * ncell * nc);
  memcpy(I->phaseFunctionIndex, pfIndex, sizeof(int32_t) * ncell * nc);
  /* MCRT:233-234 */
  float* last = I->cumulativeExt + (size_t)(nc - 1) * ncell;
  for (size_t i = 0; i < ncell; i++)
    if (fabsf(last[i] - 1.0f) <= f_spacing(1.0f)) last[i] = 1.0f + f_spacing(1.0f);
  I->forwardTables = (owned_table*)calloc(nc, sizeof(owned_table));
  I->fluxUp = (float*)calloc((size_t)nx * ny, sizeof(float));
  I->fluxDown = (float*)calloc((size_t)nx * ny, sizeof(float));
  I->fluxAbsorbed = (float*)calloc((size_t)nx * ny, sizeof(float));
  I->volumeAbsorption = (float*)calloc(ncell, sizeof(float));
  I->readyToCompute = 1;
  set_msg(I, "");
  *out = I;
  return I3RC_SUCCESS;
}

static void free_table(owned_table* t) {
  free(t->coef_offsets);
  free(t->coefs);
  free(t->angles);
  free(t->values);
  memset(t, 0, sizeof(*t));
}
static void free_matrix(matrix* m) {
  if (m) {
    free(m->values);
    m->values = NULL;
    m->numX = m->numY = 0;
  }
}

int orc_set_phase_table(orc_integrator* I, int comp, const orc_phase_table* t) {
  if (comp < 0 || comp >= I->nc) return I3RC_FAILURE;
  owned_table* o = &I->forwardTables[comp];
  free_table(o);
  o->kind = t->kind;
  o->n_entries = t->n_entries;
  if (t->kind == 1) {
    o->coef_offsets = (int32_t*)malloc(sizeof(int32_t) * (t->n_entries + 1));
    memcpy(o->coef_offsets, t->coef_offsets, sizeof(int32_t) * (t->n_entries + 1));
    o->coefs = dupf(t->coefs, t->coef_offsets[t->n_entries]);
  } else {
    o->n_angles = t->n_angles;
    o->angles = dupf(t->angles, t->n_angles);
    o->values = dupf(t->values, (size_t)t->n_angles * t->n_entries);
    /* tabulated phase functions are normalised by their constructor (SPF:322-325) and again by
       the copy that getOpticalPropertiesByComponent hands to the integrator (SPF:433, OPT:525) */
    for (int rep = 0; rep < 2; rep++)
      for (int e = 0; e < t->n_entries; e++) {
        float* v = o->values + (size_t)e * t->n_angles;
        float* tmp = dupf(v, t->n_angles);
        orc_normalize_phase_function(o->angles, tmp, t->n_angles, v);
        free(tmp);
      }
  }
  if (I->inversePhaseFunctions) free_matrix(&I->inversePhaseFunctions[comp]);
  if (I->tabulatedPhaseFunctions) free_matrix(&I->tabulatedPhaseFunctions[comp]);
  if (I->tabulatedOrigPhaseFunctions) free_matrix(&I->tabulatedOrigPhaseFunctions[comp]);
  return I3RC_SUCCESS;
}

int orc_new_Integrator_components(int nx, int ny, int nz, const float* xPos, const float* yPos, const float* zPos,
                                  int nc, const orc_component* comps, orc_integrator** out) {
  size_t ncell = (size_t)nx * ny * nz;
  float* te = (float*)malloc(sizeof(float) * ncell);
  float* ce = (float*)malloc(sizeof(float) * ncell * nc);
  float* sa = (float*)malloc(sizeof(float) * ncell * nc);
  int32_t* pi = (int32_t*)malloc(sizeof(int32_t) * ncell * nc);
  int rc = orc_getOpticalPropertiesByComponent(nx, ny, nz, nc, comps, te, ce, sa, pi);
  if (rc == I3RC_SUCCESS) rc = orc_new_Integrator(nx, ny, nz, nc, xPos, yPos, zPos, te, ce, sa, pi, out);
  if (rc == I3RC_SUCCESS)
    for (int c = 0; c < nc; c++) orc_set_phase_table(*out, c, &comps[c].table);
  free(te);
  free(ce);
  free(sa);
  free(pi);
  return rc;
}

static owned_table copy_table(const owned_table* t) {
  owned_table o = *t;
  if (t->kind == 1) {
    o.coef_offsets = (int32_t*)malloc(sizeof(int32_t) * (t->n_entries + 1));
    memcpy(o.coef_offsets, t->coef_offsets, sizeof(int32_t) * (t->n_entries + 1));
    o.coefs = dupf(t->coefs, t->coef_offsets[t->n_entries]);
  } else if (t->kind == 2) {
    o.angles = dupf(t->angles, t->n_angles);
    o.values = dupf(t->values, (size_t)t->n_angles * t->n_entries);
  }
  return o;
}
static matrix copy_matrix(const matrix* m) {
  matrix o = *m;
  if (m->values) o.values = dupf(m->values, (size_t)m->numX * m->numY);
  return o;
}

int orc_copy_Integrator(const orc_integrator* S, orc_integrator** out) { /* MCRT:1082-1253 (a full copy) */
  orc_integrator* I = (orc_integrator*)malloc(sizeof(orc_integrator));
  *I = *S;
  size_t ncell = (size_t)S->nx * S->ny * S->nz, ncol = (size_t)S->nx * S->ny;
  I->xPosition = edges1(S->xPosition + 1, S->nx + 1);
  I->yPosition = edges1(S->yPosition + 1, S->ny + 1);
  I->zPosition = edges1(S->zPosition + 1, S->nz + 1);
  I->totalExt = dupf(S->totalExt, ncell);
  I->cumulativeExt = dupf(S->cumulativeExt, ncell * S->nc);
  I->ssa = dupf(S->ssa, ncell * S->nc);
  I->phaseFunctionIndex = (int32_t*)malloc(sizeof(int32_t) * ncell * S->nc);
  memcpy(I->phaseFunctionIndex, S->phaseFunctionIndex, sizeof(int32_t) * ncell * S->nc);
  if (S->useSurfaceBDRF || S->surf_x) {
    I->surf_x = edges1(S->surf_x + 1, S->surf_nx + 1);
    I->surf_y = edges1(S->surf_y + 1, S->surf_ny + 1);
    I->surf_params = dupf(S->surf_params, (size_t)S->surf_nx * S->surf_ny);
  }
  I->forwardTables = (owned_table*)calloc(S->nc, sizeof(owned_table));
  for (int c = 0; c < S->nc; c++) I->forwardTables[c] = copy_table(&S->forwardTables[c]);
#define CPM(name)                                                     \
  if (S->name) {                                                      \
    I->name = (matrix*)calloc(S->nc, sizeof(matrix));                 \
    for (int c = 0; c < S->nc; c++) I->name[c] = copy_matrix(&S->name[c]); \
  }
  CPM(tabulatedPhaseFunctions)
  CPM(tabulatedOrigPhaseFunctions)
  CPM(inversePhaseFunctions)
  if (S->intensityDirections) I->intensityDirections = dupf(S->intensityDirections, (size_t)3 * S->nDir);
  if (S->intensityExcess) I->intensityExcess = dupf(S->intensityExcess, (size_t)(S->nc + 1) * S->nDir);
  I->fluxUp = dupf(S->fluxUp, ncol);
  I->fluxDown = dupf(S->fluxDown, ncol);
  I->fluxAbsorbed = dupf(S->fluxAbsorbed, ncol);
  I->volumeAbsorption = dupf(S->volumeAbsorption, ncell);
  if (S->intensity) I->intensity = dupf(S->intensity, ncol * S->nDir);
  if (S->intensityByComponent) I->intensityByComponent = dupf(S->intensityByComponent, ncol * S->nDir * (S->nc + 1));
  *out = I;
  return I3RC_SUCCESS;
}

void orc_finalize_Integrator(orc_integrator* I) { /* MCRT:1258-1349 */
  if (!I) return;
  free(I->xPosition);
  free(I->yPosition);
  free(I->zPosition);
  free(I->totalExt);
  free(I->cumulativeExt);
  free(I->ssa);
  free(I->phaseFunctionIndex);
  free(I->surf_x);
  free(I->surf_y);
  free(I->surf_params);
  for (int c = 0; c < I->nc; c++) {
    free_table(&I->forwardTables[c]);
    if (I->tabulatedPhaseFunctions) free_matrix(&I->tabulatedPhaseFunctions[c]);
    if (I->tabulatedOrigPhaseFunctions) free_matrix(&I->tabulatedOrigPhaseFunctions[c]);
    if (I->inversePhaseFunctions) free_matrix(&I->inversePhaseFunctions[c]);
  }
  free(I->forwardTables);
  free(I->tabulatedPhaseFunctions);
  free(I->tabulatedOrigPhaseFunctions);
  free(I->inversePhaseFunctions);
  free(I->intensityDirections);
  free(I->intensityExcess);
  free(I->fluxUp);
  free(I->fluxDown);
  free(I->fluxAbsorbed);
  free(I->volumeAbsorption);
  free(I->intensity);
  free(I->intensityByComponent);
  free(I);
}
int orc_isReady_Integrator(const orc_integrator* h) { return h && h->readyToCompute; }

/* MCRT:2041-2059 */
static void makeDirectionCosines(float mu, float phi, float* S) {
  float sinTheta = sqrtf(1.0f - mu * mu);
  float cosPhi = cosf(phi);
  float sinPhi = sinf(phi);
  S[0] = sinTheta * cosPhi;
  S[1] = sinTheta * sinPhi;
  S[2] = mu;
}
/* MCRT:2063-2082 */
static float makePeriodic(float a, float aMin, float aMax) {
  float r = a;
  for (;;) {
    if (r <= aMax && r > aMin) break;
    if (r > aMax)
      r = r - (aMax - aMin);
    else if (r == aMin)
      r = aMax;
    else
      r = r + (aMax - aMin);
  }
  return r;
}

/* MCRT:830-1069 specifyParameters */
int orc_specifyParameters(orc_integrator* I, const orc_params* p) {
  uint32_t m = p->present;
  int warn = 0, fail = 0;
#define HAS(b) ((m & (b)) != 0)
  if (HAS(I3RC_P_surfaceBDRF) && HAS(I3RC_P_surfaceAlbedo)) {
    set_msg(I, "specifyParameters: only one surface specification can be provided");
    fail = 1;
  }
  if (HAS(I3RC_P_surfaceAlbedo) && (p->surfaceAlbedo > 1.0f || p->surfaceAlbedo < 0.0f)) {
    set_msg(I, "specifyParameters: surface albedo out of range.");
    fail = 1;
  }
  if (HAS(I3RC_P_surfaceBDRF) && !(p->surf_x && p->surf_y && p->surf_params && p->surf_nx > 0 && p->surf_ny > 0)) {
    set_msg(I, "specifyParameters: surface description isn't valid.");
    fail = 1;
  }
  if (HAS(I3RC_P_minForwardTableSize) && p->minForwardTableSize < 9001) {
    set_msg(I, "specifyParameters: minForwardTableSize less than default. Value ignored.");
    warn = 1;
  }
  if (HAS(I3RC_P_minInverseTableSize) && p->minInverseTableSize < 9001) {
    set_msg(I, "specifyParameters: minInverseTableSize less than default. Value ignored.");
    warn = 1;
  }
  if (HAS(I3RC_P_hybridPhaseFunWidth) && (p->hybridPhaseFunWidth > 30.0f || p->hybridPhaseFunWidth < 0.0f)) {
    set_msg(I, "specifyParameters: hybridPhaseFunWidth out of range (0 to 30degrees).Using default (7)");
    warn = 1;
  }
  if (HAS(I3RC_P_numOrdersOrigPhaseFunIntenCalcs) && p->numOrdersOrigPhaseFunIntenCalcs < 0) {
    set_msg(I, "specifyParameters: numOrdersExactPhaseFunIntenCalcs less than 0.Using default (0)");
    warn = 1;
  }
  if (HAS(I3RC_P_maxIntensityContribution) && p->maxIntensityContribution <= 0.0f) {
    set_msg(I, "specifyParameters: maxIntensityContribution <= 0. Value is unchanged.");
    warn = 1;
  }
  if (HAS(I3RC_P_intensityMus) != HAS(I3RC_P_intensityPhis)) {
    set_msg(I, "specifyParameters: Both or neither of intensityMus and intensityPhis must be supplied");
    fail = 1;
  }
  if (HAS(I3RC_P_intensityMus) && !fail) {
    for (int i = 0; i < p->numIntensityDirections; i++) {
      if (p->intensityMus[i] < -1.0f || p->intensityMus[i] > 1.0f) {
        set_msg(I, "specifyParameters: intensityMus must be between -1 and 1");
        fail = 1;
      }
      if (fabsf(p->intensityMus[i]) < F_TINY) {
        set_msg(I, "specifyParameters: intensityMus can't be 0 (directly sideways)");
        fail = 1;
      }
      if (p->intensityPhis[i] < 0.0f || p->intensityPhis[i] > 360.0f) {
        set_msg(I, "specifyParameters: intensityPhis must be between 0 and 360");
        fail = 1;
      }
    }
  }
  if (HAS(I3RC_P_computeIntensity)) {
    if (!p->computeIntensity && HAS(I3RC_P_intensityMus)) {
      set_msg(I, "specifyParameters: intensity directions *and* computeIntensity set to false.Will compute intensity at given angles.");
      warn = 1;
    }
    if (p->computeIntensity && !HAS(I3RC_P_intensityMus) && !I->intensityDirections) {
      set_msg(I, "specifyParameters: Can't compute intensity without specifying directions.");
      fail = 1;
    }
  }
  if (fail) return I3RC_FAILURE;

  if (HAS(I3RC_P_surfaceAlbedo)) {
    I->surfaceAlbedo = p->surfaceAlbedo;
    I->useSurfaceBDRF = 0;
  } else if (HAS(I3RC_P_surfaceBDRF)) {
    free(I->surf_x);
    free(I->surf_y);
    free(I->surf_params);
    I->surf_nx = p->surf_nx;
    I->surf_ny = p->surf_ny;
    I->surf_x = edges1(p->surf_x, p->surf_nx + 1);
    I->surf_y = edges1(p->surf_y, p->surf_ny + 1);
    I->surf_params = dupf(p->surf_params, (size_t)p->surf_nx * p->surf_ny);
    I->useSurfaceBDRF = 1;
  }
  if (HAS(I3RC_P_useRayTracing)) I->useRayTracing = p->useRayTracing != 0;
  if (HAS(I3RC_P_minForwardTableSize))
    I->minForwardTableSize = p->minForwardTableSize > 9001 ? p->minForwardTableSize : 9001;
  if (HAS(I3RC_P_minInverseTableSize))
    I->minInverseTableSize = p->minInverseTableSize > 9001 ? p->minInverseTableSize : 9001;
  if (HAS(I3RC_P_useRussianRoulette)) I->useRussianRoulette = p->useRussianRoulette != 0;
  if (HAS(I3RC_P_useRussianRouletteForIntensity))
    I->useRussianRouletteForIntensity = p->useRussianRouletteForIntensity != 0;
  if (HAS(I3RC_P_zetaMin)) {
    if (p->zetaMin < 0.0f) {
      set_msg(I, "specifyParameters: zetaMin must be >= 0. Value is unchanged.");
      warn = 1;
    } else {
      I->zetaMin = p->zetaMin;
      if (p->zetaMin > 1.0f) {
        set_msg(I, "specifyParameters: zetaMin > 1. That's kind of large.");
        warn = 1;
      }
    }
  }
  if (HAS(I3RC_P_useHybridPhaseFunsForIntenCalcs))
    I->useHybridPhaseFunsForIntenCalcs = p->useHybridPhaseFunsForIntenCalcs != 0;
  if (HAS(I3RC_P_hybridPhaseFunWidth)) {
    if (p->hybridPhaseFunWidth > 0.0f && p->hybridPhaseFunWidth < 30.0f)
      I->hybridPhaseFunWidth = p->hybridPhaseFunWidth;
    else
      I->hybridPhaseFunWidth = 7.0f;
    /* MCRT:999-1004 (quirk Q7): the forward tables are dropped so they are re-tabulated */
    if (I->tabulatedPhaseFunctions) {
      for (int c = 0; c < I->nc; c++) free_matrix(&I->tabulatedPhaseFunctions[c]);
      free(I->tabulatedPhaseFunctions);
      I->tabulatedPhaseFunctions = NULL;
    }
  }
  if (HAS(I3RC_P_numOrdersOrigPhaseFunIntenCalcs))
    I->numOrdersOrigPhaseFunIntenCalcs = p->numOrdersOrigPhaseFunIntenCalcs >= 0 ? p->numOrdersOrigPhaseFunIntenCalcs : 0;
  if (HAS(I3RC_P_limitIntensityContributions)) I->limitIntensityContributions = p->limitIntensityContributions != 0;
  if (HAS(I3RC_P_maxIntensityContribution) && p->maxIntensityContribution > 0.0f)
    I->maxIntensityContribution = p->maxIntensityContribution;
  size_t ncol = (size_t)I->nx * I->ny;
  if (HAS(I3RC_P_intensityMus)) {
    free(I->intensityDirections);
    free(I->intensity);
    free(I->intensityByComponent);
    I->nDir = p->numIntensityDirections;
    I->intensityDirections = (float*)malloc(sizeof(float) * 3 * I->nDir);
    I->intensity = (float*)calloc(ncol * I->nDir, sizeof(float));
    I->intensityByComponent = (float*)calloc(ncol * I->nDir * (I->nc + 1), sizeof(float));
    for (int i = 0; i < I->nDir; i++)
      makeDirectionCosines(p->intensityMus[i], p->intensityPhis[i] * Pi / 180.0f, I->intensityDirections + 3 * i);
    I->computeIntensity = 1;
  }
  if (HAS(I3RC_P_computeIntensity)) {
    if (!p->computeIntensity && !HAS(I3RC_P_intensityMus)) {
      free(I->intensityDirections);
      free(I->intensity);
      free(I->intensityByComponent);
      I->intensityDirections = I->intensity = I->intensityByComponent = NULL;
      I->nDir = 0;
      I->computeIntensity = 0;
    }
  }
  if (I->computeIntensity && I->limitIntensityContributions) {
    free(I->intensityExcess);
    I->intensityExcess = (float*)calloc((size_t)(I->nc + 1) * I->nDir, sizeof(float));
  }
  if (!warn) set_msg(I, "");
  return warn ? I3RC_WARNING : I3RC_SUCCESS;
}

/* MCRT:1809-1923 */
static int tabulateInversePhaseFunctions(orc_integrator* I) {
  if (!I->inversePhaseFunctions) I->inversePhaseFunctions = (matrix*)calloc(I->nc, sizeof(matrix));
  for (int c = 0; c < I->nc; c++) {
    matrix* M = &I->inversePhaseFunctions[c];
    if (M->values && M->numX >= I->minInverseTableSize) continue;
    owned_table* ot = &I->forwardTables[c];
    if (ot->kind == 0) {
      set_msg(I, "tabulateInversePhaseFunctions: failed on component");
      return I3RC_FAILURE;
    }
    orc_phase_table t = {ot->kind, ot->n_entries, ot->coef_offsets, ot->coefs, ot->n_angles, ot->angles, ot->values};
    int nSteps = I->minInverseTableSize;
    free_matrix(M);
    M->numX = nSteps;
    M->numY = ot->n_entries;
    M->values = (float*)malloc(sizeof(float) * (size_t)nSteps * ot->n_entries);
    for (int e = 0; e < ot->n_entries; e++) orc_inverse_phase_function(&t, e, nSteps, M->values + (size_t)e * nSteps);
  }
  return I3RC_SUCCESS;
}
static int tabulateForwardPhaseFunctions(orc_integrator* I) {
  if (!I->tabulatedPhaseFunctions) I->tabulatedPhaseFunctions = (matrix*)calloc(I->nc, sizeof(matrix));
  if (!I->tabulatedOrigPhaseFunctions) I->tabulatedOrigPhaseFunctions = (matrix*)calloc(I->nc, sizeof(matrix));
  for (int c = 0; c < I->nc; c++) {
    matrix* M = &I->tabulatedPhaseFunctions[c];
    if (M->values && M->numX >= I->minForwardTableSize) continue;
    owned_table* ot = &I->forwardTables[c];
    if (ot->kind == 0) {
      set_msg(I, "tabulatePhaseFunctions: failed on component");
      return I3RC_FAILURE;
    }
    orc_phase_table t = {ot->kind, ot->n_entries, ot->coef_offsets, ot->coefs, ot->n_angles, ot->angles, ot->values};
    int nSteps = I->minForwardTableSize;
    float* angles = (float*)malloc(sizeof(float) * nSteps);
    for (int j = 0; j < nSteps; j++) angles[j] = (float)j / (float)(nSteps - 1) * Pi; /* MCRT:1900 */
    float* temp = (float*)malloc(sizeof(float) * (size_t)nSteps * ot->n_entries);
    orc_phase_values_table(&t, angles, nSteps, temp);
    matrix* O = &I->tabulatedOrigPhaseFunctions[c];
    free_matrix(O);
    O->numX = nSteps;
    O->numY = ot->n_entries;
    O->values = dupf(temp, (size_t)nSteps * ot->n_entries);
    free_matrix(M);
    M->numX = nSteps;
    M->numY = ot->n_entries;
    M->values = (float*)malloc(sizeof(float) * (size_t)nSteps * ot->n_entries);
    if (I->useHybridPhaseFunsForIntenCalcs && I->hybridPhaseFunWidth > 0.0f)
      orc_hybrid_phase_functions(angles, nSteps, ot->n_entries, temp, I->hybridPhaseFunWidth, M->values);
    else
      memcpy(M->values, temp, sizeof(float) * (size_t)nSteps * ot->n_entries);
    free(angles);
    free(temp);
  }
  return I3RC_SUCCESS;
}
int orc_tabulate(orc_integrator* I) {
  int rc = tabulateInversePhaseFunctions(I);
  if (rc == I3RC_SUCCESS && I->computeIntensity) rc = tabulateForwardPhaseFunctions(I);
  return rc;
}
int orc_get_table(orc_integrator* I, int which, int comp, float* out, int* nSteps, int* nEntries) {
  matrix* arr = which == 0 ? I->inversePhaseFunctions : which == 1 ? I->tabulatedPhaseFunctions : I->tabulatedOrigPhaseFunctions;
  if (!arr || comp < 0 || comp >= I->nc || !arr[comp].values) return I3RC_FAILURE;
  if (nSteps) *nSteps = arr[comp].numX;
  if (nEntries) *nEntries = arr[comp].numY;
  if (out) memcpy(out, arr[comp].values, sizeof(float) * (size_t)arr[comp].numX * arr[comp].numY);
  return I3RC_SUCCESS;
}

/* MCRT:1353-1388 */
static void findXYIndicies(const orc_integrator* I, float xPos, float yPos, int* xIndex, int* yIndex) {
  if (I->xyRegularlySpaced) {
    int xi = (int)((xPos - I->x0) / I->deltaX) + 1;
    int yi = (int)((yPos - I->y0) / I->deltaY) + 1;
    if (xi > I->nx) xi = I->nx;
    if (yi > I->ny) yi = I->ny;
    if (fabsf(I->xPosition[xi + 1] - xPos) < f_spacing(xPos)) xi = xi + 1;
    if (fabsf(I->yPosition[yi + 1] - yPos) < f_spacing(yPos)) yi = yi + 1;
    if (xi == I->nx + 1) xi = 1;
    if (yi == I->ny + 1) yi = 1;
    *xIndex = xi;
    *yIndex = yi;
  } else {
    *xIndex = findIndex1(xPos, I->xPosition, I->nx + 1, *xIndex);
    *yIndex = findIndex1(yPos, I->yPosition, I->ny + 1, *yIndex);
  }
}
static void findZIndex(const orc_integrator* I, float zPos, int* zIndex) {
  if (I->zRegularlySpaced) {
    int zi = (int)((zPos - I->z0) / I->deltaZ) + 1;
    if (zi > I->nz) zi = I->nz;
    if (fabsf(I->zPosition[zi + 1] - zPos) < f_spacing(zPos)) zi = zi + 1;
    *zIndex = zi;
  } else {
    *zIndex = findIndex1(zPos, I->zPosition, I->nz + 1, *zIndex);
  }
}

/* MCRT:1390-1417 (quirk Q1 kept: leftOver is not scaled by numIntervals) */
static float computeScatteringAngle(float randomDeviate, const float* table0, int numIntervals) {
  const float* T = table0 - 1;
  int angleIndex = (int)(randomDeviate * (float)numIntervals) + 1;
  if (angleIndex < numIntervals) {
    float leftOver = randomDeviate - (float)(angleIndex - 1) / (float)numIntervals;
    return (1.0f - leftOver) * T[angleIndex] + leftOver * T[angleIndex + 1];
  }
  return T[numIntervals];
}

/* MCRT:1613-1652 */
static float lookUpPhaseFuncValFromTable(const float* table0, int nAngleSteps, float scatteringAngle) {
  const float* T = table0 - 1;
  float deltaTheta = Pi / (float)(nAngleSteps - 1);
  int idx = (int)(scatteringAngle / deltaTheta) + 1;
  if (idx < nAngleSteps) {
    float w = 1.0f - (scatteringAngle - (float)(idx - 1) * deltaTheta) / deltaTheta;
    return w * T[idx] + (1.0f - w) * T[idx + 1];
  }
  return T[nAngleSteps];
}

/* MCRT:1654-1807 accumulateExtinctionAlongPath; hasLimit mirrors present(extToAccumulate) */
static void accumulateExtinctionAlongPath(orc_integrator* I, const float* dc, float* xPos, float* yPos, float* zPos,
                                          int* xIndex, int* yIndex, int* zIndex, float* extAccumulated, int hasLimit,
                                          float extToAccumulate, int64_t* crossings) {
  int nXcells = I->nx, nYcells = I->ny, nZcells = I->nz;
  int SideIncrement[3], CellIncrement[3];
  float step[3];
  float ext = 0.0f, totalPath = 0.0f;
  for (int a = 0; a < 3; a++) {
    SideIncrement[a] = dc[a] >= 0.0f ? 1 : 0;
    CellIncrement[a] = dc[a] >= 0.0f ? 1 : -1;
  }
  float z0 = I->zPosition[1], zMax = I->zPosition[nZcells + 1];
  for (;;) {
    if (fabsf(dc[0]) >= 2.0f * F_TINY)
      step[0] = (I->xPosition[*xIndex + SideIncrement[0]] - *xPos) / dc[0];
    else
      step[0] = F_HUGE;
    if (fabsf(dc[1]) >= 2.0f * F_TINY)
      step[1] = (I->yPosition[*yIndex + SideIncrement[1]] - *yPos) / dc[1];
    else
      step[1] = F_HUGE;
    if (fabsf(dc[2]) >= 2.0f * F_TINY)
      step[2] = (I->zPosition[*zIndex + SideIncrement[2]] - *zPos) / dc[2];
    else
      step[2] = F_HUGE;
    float thisStep = step[0] < step[1] ? step[0] : step[1];
    if (step[2] < thisStep) thisStep = step[2];
    if (thisStep <= 0.0f) {
      ext = -2.0f;
      break;
    }
    float thisCellExt = TE(I, *xIndex, *yIndex, *zIndex);
    (*crossings)++;
    if (hasLimit) {
      if (ext + thisStep * thisCellExt > extToAccumulate) {
        thisStep = (extToAccumulate - ext) / thisCellExt;
        *xPos = *xPos + thisStep * dc[0];
        *yPos = *yPos + thisStep * dc[1];
        *zPos = *zPos + thisStep * dc[2];
        totalPath = totalPath + thisStep;
        ext = extToAccumulate;
        break;
      }
    }
    ext = ext + thisStep * thisCellExt;
    totalPath = totalPath + thisStep;

    if (step[0] <= thisStep) {
      *xPos = I->xPosition[*xIndex + SideIncrement[0]];
      *xIndex = *xIndex + CellIncrement[0];
    } else {
      *xPos = *xPos + thisStep * dc[0];
      if (fabsf(I->xPosition[*xIndex + SideIncrement[0]] - *xPos) <= 2.0f * f_spacing(*xPos))
        *xIndex = *xIndex + CellIncrement[0];
    }
    if (step[1] <= thisStep) {
      *yPos = I->yPosition[*yIndex + SideIncrement[1]];
      *yIndex = *yIndex + CellIncrement[1];
    } else {
      *yPos = *yPos + thisStep * dc[1];
      if (fabsf(I->yPosition[*yIndex + SideIncrement[1]] - *yPos) <= 2.0f * f_spacing(*yPos))
        *yIndex = *yIndex + CellIncrement[1];
    }
    if (step[2] <= thisStep) {
      *zPos = I->zPosition[*zIndex + SideIncrement[2]];
      *zIndex = *zIndex + CellIncrement[2];
    } else {
      *zPos = *zPos + thisStep * dc[2];
      if (fabsf(I->zPosition[*zIndex + SideIncrement[2]] - *zPos) <= 2.0f * f_spacing(*zPos))
        *zIndex = *zIndex + CellIncrement[2];
    }
    /* periodicity; quirk Q2: the y nudge uses cellIncrement(1) (MCRT:1784,1787) */
    if (*xIndex <= 0) {
      *xIndex = nXcells;
      *xPos = I->xPosition[*xIndex + 1] + (float)(CellIncrement[0] * 2) * f_spacing(*xPos);
    } else if (*xIndex >= nXcells + 1) {
      *xIndex = 1;
      *xPos = I->xPosition[*xIndex] + (float)(CellIncrement[0] * 2) * f_spacing(*xPos);
    }
    if (*yIndex <= 0) {
      *yIndex = nYcells;
      *yPos = I->yPosition[*yIndex + 1] + (float)(CellIncrement[0] * 2) * f_spacing(*yPos);
    } else if (*yIndex >= nYcells + 1) {
      *yIndex = 1;
      *yPos = I->yPosition[*yIndex] + (float)(CellIncrement[0] * 2) * f_spacing(*yPos);
    }
    if (*zIndex > nZcells) {
      *zPos = zMax + 2.0f * f_spacing(zMax);
      break;
    }
    if (*zIndex < 1) {
      *zPos = z0;
      break;
    }
  }
  (void)totalPath;
  *extAccumulated = ext;
}

/* MCRT:2086-2113 NEXT_DIRECT */
static void next_direct(uint32_t* rng, int64_t* draws, float scatteringCosine, float* S) {
  float D = 2.0f, AX = 0.0f, AY = 0.0f, B;
  while (D > 1.0f) {
    AX = 1.0f - 2.0f * orc_mt_real(rng);
    AY = 1.0f - 2.0f * orc_mt_real(rng);
    *draws += 2;
    D = AX * AX + AY * AY;
  }
  B = sqrtf((1.0f - scatteringCosine * scatteringCosine) / D);
  AX = AX * B;
  AY = AY * B;
  B = S[0] * AX - S[1] * AY;
  D = scatteringCosine - B / (1.0f + fabsf(S[2]));
  S[0] = S[0] * D + AX;
  S[1] = S[1] * D - AY;
  S[2] = S[2] * scatteringCosine - f_sign(B, S[2] * B);
}
void orc_next_direct(const float* xi, int nxi, float scatteringCosine, float* S) {
  float D = 2.0f, AX = 0.0f, AY = 0.0f, B;
  int k = 0;
  while (D > 1.0f && k + 1 < nxi + 1) {
    AX = 1.0f - 2.0f * xi[k];
    AY = 1.0f - 2.0f * xi[k + 1];
    k += 2;
    D = AX * AX + AY * AY;
    if (k >= nxi) break;
  }
  B = sqrtf((1.0f - scatteringCosine * scatteringCosine) / D);
  AX = AX * B;
  AY = AY * B;
  B = S[0] * AX - S[1] * AY;
  D = scatteringCosine - B / (1.0f + fabsf(S[2]));
  S[0] = S[0] * D + AX;
  S[1] = S[1] * D - AY;
  S[2] = S[2] * scatteringCosine - f_sign(B, S[2] * B);
}

/* SURF:121-162 */
static float computeSurfaceReflectance(const orc_integrator* I, float xPos, float yPos) {
  float x0 = I->surf_x[1], y0 = I->surf_y[1];
  float xMax = I->surf_x[I->surf_nx + 1], yMax = I->surf_y[I->surf_ny + 1];
  int xi = findIndex1(makePeriodic(xPos, x0, xMax), I->surf_x, I->surf_nx + 1, 0);
  int yi = findIndex1(makePeriodic(yPos, y0, yMax), I->surf_y, I->surf_ny + 1, 0);
  if (xi < 1) xi = 1;
  if (yi < 1) yi = 1;
  if (xi > I->surf_nx) xi = I->surf_nx;
  if (yi > I->surf_ny) yi = I->surf_ny;
  return I->surf_params[(size_t)(yi - 1) * I->surf_nx + (xi - 1)];
}

/* MCRT:1419-1611 computeIntensityContribution */
static void computeIntensityContribution(orc_integrator* I, float photonWeight, float xPos, float yPos, float zPos,
                                         int xIndex, int yIndex, int zIndex, const float* dc, int component,
                                         uint32_t* rng, int scatteringOrder, float* contributions, int* xIndexF,
                                         int* yIndexF) {
  int nD = I->nDir;
  int zIndexMax = I->nz + 1;
  float normalizedPhaseFunc[64], tausToBoundary[64];
  int zIndexF[64];
  for (int i = 0; i < nD; i++) {
    xIndexF[i] = xIndex;
    yIndexF[i] = yIndex;
    zIndexF[i] = zIndex;
  }
  if (component < 1) {
    for (int i = 0; i < nD; i++) normalizedPhaseFunc[i] = 1.0f / Pi; /* quirk Q10 */
  } else {
    int pfi = F4(I, phaseFunctionIndex, xIndex, yIndex, zIndex, component);
    const matrix* M;
    if (I->useHybridPhaseFunsForIntenCalcs && scatteringOrder <= I->numOrdersOrigPhaseFunIntenCalcs)
      M = &I->tabulatedOrigPhaseFunctions[component - 1];
    else
      M = &I->tabulatedPhaseFunctions[component - 1];
    const float* col = M->values + (size_t)(pfi - 1) * M->numX;
    for (int i = 0; i < nD; i++) {
      const float* d = I->intensityDirections + 3 * i;
      float proj = dc[0] * d[0] + dc[1] * d[1] + dc[2] * d[2];
      if (fabsf(proj) > 1.0f) proj = f_sign(1.0f, proj);
      float ang = acosf(proj);
      float val = lookUpPhaseFuncValFromTable(col, M->numX, ang);
      normalizedPhaseFunc[i] = val / (4.0f * Pi * fabsf(d[2]));
    }
  }
  if (!I->useRussianRouletteForIntensity) {
    for (int i = 0; i < nD; i++) {
      float xT = xPos, yT = yPos, zT = zPos;
      accumulateExtinctionAlongPath(I, I->intensityDirections + 3 * i, &xT, &yT, &zT, &xIndexF[i], &yIndexF[i],
                                    &zIndexF[i], &tausToBoundary[i], 0, 0.0f, &I->cnt.crossings_intensity);
      if (tausToBoundary[i] >= 0.0f)
        contributions[i] = photonWeight * normalizedPhaseFunc[i] * expf(-tausToBoundary[i]);
      else
        contributions[i] = 0.0f;
    }
  } else {
    for (int i = 0; i < nD; i++) {
      float xT = xPos, yT = yPos, zT = zPos;
      const float* d = I->intensityDirections + 3 * i;
      float tauFree = -logf(fmaxf(F_TINY, orc_mt_real(rng)));
      I->cnt.rng_draws++;
      if (Pi * normalizedPhaseFunc[i] <= I->zetaMin) {
        accumulateExtinctionAlongPath(I, d, &xT, &yT, &zT, &xIndexF[i], &yIndexF[i], &zIndexF[i], &tausToBoundary[i],
                                      1, tauFree, &I->cnt.crossings_intensity);
        float xi = orc_mt_real(rng);
        I->cnt.rng_draws++;
        if (xi <= Pi * normalizedPhaseFunc[i] / I->zetaMin && zIndexF[i] >= zIndexMax)
          contributions[i] = photonWeight * I->zetaMin / Pi;
        else
          contributions[i] = 0.0f;
      } else {
        float tauMax = -logf(I->zetaMin / fmaxf(F_TINY, Pi * normalizedPhaseFunc[i]));
        accumulateExtinctionAlongPath(I, d, &xT, &yT, &zT, &xIndexF[i], &yIndexF[i], &zIndexF[i], &tausToBoundary[i],
                                      1, tauMax, &I->cnt.crossings_intensity);
        if (zIndexF[i] >= zIndexMax && tausToBoundary[i] >= 0.0f) {
          contributions[i] = photonWeight * normalizedPhaseFunc[i] * expf(-tausToBoundary[i]);
        } else if (tausToBoundary[i] >= 0.0f) {
          /* quirk Q4: a ray that left through the bottom has zIndexF = 0; the reference would read
             zPosition(0).  The oracle gives such rays zero contribution instead of reading out of bounds. */
          if (zIndexF[i] < 1) {
            contributions[i] = 0.0f;
          } else {
            accumulateExtinctionAlongPath(I, d, &xT, &yT, &zT, &xIndexF[i], &yIndexF[i], &zIndexF[i],
                                          &tausToBoundary[i], 1, tauFree, &I->cnt.crossings_intensity);
            if (zIndexF[i] >= zIndexMax)
              contributions[i] = photonWeight * I->zetaMin / Pi;
            else
              contributions[i] = 0.0f;
          }
        } else {
          contributions[i] = 0.0f;
        }
      }
    }
  }
  if (I->limitIntensityContributions) {
    for (int i = 0; i < nD; i++)
      if (contributions[i] > I->maxIntensityContribution) {
        I->intensityExcess[(size_t)component * nD + i] += contributions[i] - I->maxIntensityContribution;
        contributions[i] = I->maxIntensityContribution;
      }
  }
  /* a ray that left through the bottom keeps xIndexF/yIndexF inside 1..n, so the caller's tally is safe */
}

/* photon stream (ILL:34-41) */
typedef struct {
  int64_t n, current;
  float *x, *y, *z, *mu, *phi;
} stream;

static int new_PhotonStream(const orc_photon_source* s, uint32_t* rng, stream* ph, orc_integrator* I) {
  int64_t n = s->numberOfPhotons;
  if (n <= 0) {
    set_msg(I, "setIllumination: must ask for non-negative number of photons.");
    return I3RC_FAILURE;
  }
  float pi = acosf(-1.0f);
  switch (s->kind) {
    case I3RC_SRC_DIRECTIONAL:
    case I3RC_SRC_SPOTLIGHT:
      if (s->solarAzimuth < 0.0f || s->solarAzimuth > 360.0f) {
        set_msg(I, "setIllumination: solarAzimuth out of bounds");
        return I3RC_FAILURE;
      }
      /* fall through */
    case I3RC_SRC_RANDOM_AZIMUTH:
      if (fabsf(s->solarMu) > 1.0f || fabsf(s->solarMu) <= F_TINY) {
        set_msg(I, "setIllumination: solarMu out of bounds");
        return I3RC_FAILURE;
      }
      break;
    default:
      break;
  }
  if (s->kind == I3RC_SRC_SPOTLIGHT && (s->x > 1.0f || s->x <= 0.0f || s->y > 1.0f || s->y <= 0.0f)) {
    set_msg(I, "setIllumination: x and y positions must be between 0 and 1");
    return I3RC_FAILURE;
  }
  if (s->kind == I3RC_SRC_INTERNAL_FLUX || s->kind == I3RC_SRC_INTERNAL_INTENSITY) {
    if (s->x > 1.0f || s->x <= 0.0f || s->y > 1.0f || s->y <= 0.0f || s->z > 1.0f || s->z <= 0.0f) {
      set_msg(I, "setIllumination: x, y, z positions must be between 0 and 1");
      return I3RC_FAILURE;
    }
    if ((s->has_deltaX && (s->x + s->deltaX / 2.0f > 1.0f || s->x - s->deltaX / 2.0f <= 0.0f)) ||
        (s->has_deltaY && (s->y + s->deltaY / 2.0f > 1.0f || s->y - s->deltaY / 2.0f <= 0.0f))) {
      set_msg(I, "setIllumination: max, min positions must be between 0 and 1");
      return I3RC_FAILURE;
    }
  }
  if (s->kind == I3RC_SRC_INTERNAL_INTENSITY) {
    if (s->detectorPhi < 0.0f || s->detectorPhi > 360.0f) {
      set_msg(I, "setIllumination: detectorPhi out of bounds");
      return I3RC_FAILURE;
    }
    if (fabsf(s->detectorMu) > 1.0f || fabsf(s->detectorMu) <= F_TINY) {
      set_msg(I, "setIllumination: detectorMu out of bounds");
      return I3RC_FAILURE;
    }
  }
  ph->n = n;
  ph->current = 1;
  ph->x = (float*)malloc(sizeof(float) * n);
  ph->y = (float*)malloc(sizeof(float) * n);
  ph->z = (float*)malloc(sizeof(float) * n);
  ph->mu = (float*)malloc(sizeof(float) * n);
  ph->phi = (float*)malloc(sizeof(float) * n);
  float ztop = 1.0f - f_spacing(1.0f);
  switch (s->kind) {
    case I3RC_SRC_DIRECTIONAL: /* ILL:62-104 */
      for (int64_t i = 0; i < n; i++) {
        ph->x[i] = orc_mt_real(rng);
        ph->y[i] = orc_mt_real(rng);
        ph->z[i] = ztop;
        ph->mu[i] = -fabsf(s->solarMu);
        ph->phi[i] = s->solarAzimuth * pi / 180.0f;
      }
      break;
    case I3RC_SRC_RANDOM_AZIMUTH: /* ILL:106-146 */
      for (int64_t i = 0; i < n; i++) {
        ph->x[i] = orc_mt_real(rng);
        ph->y[i] = orc_mt_real(rng);
        ph->phi[i] = orc_mt_real(rng) * 2.0f * pi;
        ph->z[i] = ztop;
        ph->mu[i] = -fabsf(s->solarMu);
      }
      break;
    case I3RC_SRC_FLUX: /* ILL:148-185 */
      for (int64_t i = 0; i < n; i++) {
        ph->x[i] = orc_mt_real(rng);
        ph->y[i] = orc_mt_real(rng);
        ph->mu[i] = -sqrtf(orc_mt_real(rng));
        ph->phi[i] = orc_mt_real(rng) * 2.0f * pi;
        ph->z[i] = ztop;
      }
      break;
    case I3RC_SRC_SPOTLIGHT: /* ILL:187-226 */
      for (int64_t i = 0; i < n; i++) {
        ph->mu[i] = -fabsf(s->solarMu);
        ph->phi[i] = s->solarAzimuth * pi / 180.0f;
        ph->x[i] = s->x;
        ph->y[i] = s->y;
        ph->z[i] = ztop;
      }
      break;
    case I3RC_SRC_INTERNAL_FLUX: { /* ILL:228-327 */
      for (int64_t i = 0; i < n; i++) {
        ph->x[i] = s->x;
        ph->y[i] = s->y;
        ph->z[i] = s->detectorPointsUp ? fmaxf(s->z, 2.0f * F_TINY) : fminf(s->z, ztop);
      }
      for (int64_t i = 0; i < n; i++) {
        ph->mu[i] = sqrtf(orc_mt_real(rng));
        ph->phi[i] = orc_mt_real(rng) * 2.0f * pi;
      }
      if (!s->detectorPointsUp)
        for (int64_t i = 0; i < n; i++) ph->mu[i] = -ph->mu[i];
      for (;;) {
        int64_t nrep = 0;
        for (int64_t i = 0; i < n; i++)
          if (fabsf(ph->mu[i]) < 2.0f * F_TINY) {
            ph->mu[i] = sqrtf(orc_mt_real(rng));
            nrep++;
          }
        if (!nrep) break;
      }
      if (s->has_deltaX)
        for (int64_t i = 0; i < n; i++) ph->x[i] = ph->x[i] + s->deltaX * (1.0f - 0.5f * orc_mt_real(rng));
      if (s->has_deltaY)
        for (int64_t i = 0; i < n; i++) ph->y[i] = ph->y[i] + s->deltaY * (1.0f - 0.5f * orc_mt_real(rng));
      break;
    }
    case I3RC_SRC_INTERNAL_INTENSITY: /* ILL:329-424; detectorPhi is stored as given (degrees), as the reference does */
      for (int64_t i = 0; i < n; i++) {
        ph->x[i] = s->x;
        ph->y[i] = s->y;
        ph->z[i] = s->detectorMu > F_TINY ? fmaxf(s->z, 2.0f * F_TINY) : fminf(s->z, ztop);
        ph->mu[i] = s->detectorMu;
        ph->phi[i] = s->detectorPhi;
      }
      if (s->has_deltaX)
        for (int64_t i = 0; i < n; i++) ph->x[i] = ph->x[i] + s->deltaX * (1.0f - 0.5f * orc_mt_real(rng));
      if (s->has_deltaY)
        for (int64_t i = 0; i < n; i++) ph->y[i] = ph->y[i] + s->deltaY * (1.0f - 0.5f * orc_mt_real(rng));
      break;
    case I3RC_SRC_ARRAYS:
      memcpy(ph->x, s->xPosition, sizeof(float) * n);
      memcpy(ph->y, s->yPosition, sizeof(float) * n);
      memcpy(ph->z, s->zPosition, sizeof(float) * n);
      memcpy(ph->mu, s->initialMu, sizeof(float) * n);
      memcpy(ph->phi, s->initialPhi, sizeof(float) * n);
      break;
    default:
      set_msg(I, "new_PhotonStream: unknown source kind");
      return I3RC_FAILURE;
  }
  return I3RC_SUCCESS;
}
static void finalize_PhotonStream(stream* ph) {
  free(ph->x);
  free(ph->y);
  free(ph->z);
  free(ph->mu);
  free(ph->phi);
  memset(ph, 0, sizeof(*ph));
}

/* MCRT:400-707 computeRT */
static int computeRT(orc_integrator* I, uint32_t* rng, stream* ph, int64_t* numPhotonsProcessed) {
  float xPos, yPos, zPos, mu, phi;
  float tauToTravel, photonWeight, scatteringAngle, tauAccumulated, ssa, maxExtinction = 0.0f;
  int useRayTracing = I->useRayTracing, useMaxCrossSection = !useRayTracing, scatterThisEvent = 1;
  int xIndex, yIndex, zIndex, component, phaseFunctionIndex;
  int scatteringOrder;
  int64_t nPhotons = 0, nBad = 0;
  float dc[3];
  int nD = I->computeIntensity ? I->nDir : 0;
  float contributions[64];
  int xIndexF[64], yIndexF[64];
  size_t ncell = (size_t)I->nx * I->ny * I->nz;
  if (nD > 64) return I3RC_FAILURE;

  if (useMaxCrossSection)
    for (size_t i = 0; i < ncell; i++)
      if (I->totalExt[i] > maxExtinction) maxExtinction = I->totalExt[i];
  float x0 = I->x0, xMax = I->xPosition[I->nx + 1];
  float y0 = I->y0, yMax = I->yPosition[I->ny + 1];
  float z0 = I->z0, zMax = I->zPosition[I->nz + 1];
  float* cumTmp = (float*)malloc(sizeof(float) * (I->nc + 2)); /* 1-based (/0, cumulativeExt(:)/) */
  const size_t ncol = (size_t)I->nx * I->ny;

  for (;;) { /* photonLoop */
    if (!(ph->current > 0 && ph->current <= ph->n)) break;
    int64_t ip = ph->current - 1;
    xPos = ph->x[ip];
    yPos = ph->y[ip];
    zPos = ph->z[ip];
    mu = ph->mu[ip];
    phi = ph->phi[ip];
    ph->current++;
    scatteringOrder = 0;
    makeDirectionCosines(mu, phi, dc);
    photonWeight = 1.0f;
    nPhotons++;
    xPos = x0 + xPos * (xMax - x0);
    yPos = y0 + yPos * (yMax - y0);
    zPos = z0 + zPos * (zMax - z0);
    xIndex = 1;
    yIndex = 1;
    zIndex = 1;
    findXYIndicies(I, xPos, yPos, &xIndex, &yIndex);
    findZIndex(I, zPos, &zIndex);

    for (;;) { /* scatteringLoop */
      tauToTravel = -logf(fmaxf(F_TINY, orc_mt_real(rng)));
      I->cnt.rng_draws++;
      if (useRayTracing) {
        accumulateExtinctionAlongPath(I, dc, &xPos, &yPos, &zPos, &xIndex, &yIndex, &zIndex, &tauAccumulated, 1,
                                      tauToTravel, &I->cnt.crossings_photon);
        if (tauAccumulated < 0.0f) {
          nBad++;
          goto nextPhoton;
        }
      } else {
        xPos = makePeriodic(xPos + dc[0] * tauToTravel / maxExtinction, x0, xMax);
        yPos = makePeriodic(yPos + dc[1] * tauToTravel / maxExtinction, y0, yMax);
        zPos = zPos + dc[2] * tauToTravel / maxExtinction;
      }

      if (zPos >= zMax) {
        if (useMaxCrossSection) {
          xPos = makePeriodic(xPos - dc[0] * fabsf((zPos - zMax) / dc[2]), x0, xMax);
          yPos = makePeriodic(yPos - dc[1] * fabsf((zPos - zMax) / dc[2]), y0, yMax);
          findXYIndicies(I, xPos, yPos, &xIndex, &yIndex);
        }
        F2(I, fluxUp, xIndex, yIndex) += photonWeight;
        I->cnt.exits_top++;
        goto nextPhoton;
      } else if (zPos <= z0 + f_spacing(z0)) {
        scatteringOrder++;
        if (useMaxCrossSection) {
          xPos = makePeriodic(xPos - dc[0] * fabsf((zPos - z0) / dc[2]), x0, xMax);
          yPos = makePeriodic(yPos - dc[1] * fabsf((zPos - z0) / dc[2]), y0, yMax);
          findXYIndicies(I, xPos, yPos, &xIndex, &yIndex);
        }
        zIndex = 1;
        zPos = z0 + f_spacing(z0);
        F2(I, fluxDown, xIndex, yIndex) += photonWeight;
        I->cnt.surface_hits++;
        for (;;) {
          mu = sqrtf(orc_mt_real(rng));
          I->cnt.rng_draws++;
          if (fabsf(mu) > 2.0f * F_TINY) break;
        }
        phi = 2.0f * Pi * orc_mt_real(rng);
        I->cnt.rng_draws++;
        if (I->useSurfaceBDRF)
          photonWeight = photonWeight * computeSurfaceReflectance(I, xPos, yPos);
        else
          photonWeight = photonWeight * I->surfaceAlbedo;
        if (photonWeight <= F_TINY) goto nextPhoton;
        makeDirectionCosines(mu, phi, dc);
        if (I->computeIntensity) {
          computeIntensityContribution(I, photonWeight, xPos, yPos, zPos, xIndex, yIndex, zIndex, dc, 0, rng,
                                       scatteringOrder, contributions, xIndexF, yIndexF);
          for (int i = 0; i < nD; i++) {
            size_t col = (size_t)(yIndexF[i] - 1) * I->nx + (xIndexF[i] - 1);
            I->intensity[(size_t)i * ncol + col] += contributions[i];
            I->intensityByComponent[(size_t)i * ncol + col] += contributions[i];
            if (contributions[i] != 0.0f) I->cnt.contributions++;
          }
        }
      } else {
        if (useMaxCrossSection) {
          /* Deviation Q11 (documented in DESIGN.md): the reference moves the photon (MCRT:494-496) but never
             refreshes xIndex/yIndex/zIndex before reading totalExt (MCRT:588), so its maximum cross-section mode
             uses the entry cell's extinction everywhere.  The oracle (and the product) look the cell up. */
          findXYIndicies(I, xPos, yPos, &xIndex, &yIndex);
          findZIndex(I, zPos, &zIndex);
          scatterThisEvent = orc_mt_real(rng) < TE(I, xIndex, yIndex, zIndex) / maxExtinction;
          I->cnt.rng_draws++;
          if (!scatterThisEvent) I->cnt.null_collisions++;
        }
        if (useRayTracing || scatterThisEvent) {
          scatteringOrder++;
          if (TE(I, xIndex, yIndex, zIndex) <= 0.0f) { /* MCRT:606-632, quirk Q3 kept */
            if (xPos - I->xPosition[xIndex] <= 0.0f && dc[0] > 0.0f) {
              xPos = xPos - f_spacing(xPos);
              xIndex = xIndex - 1;
              if (xIndex <= 0) {
                xIndex = I->nx;
                xPos = I->xPosition[xIndex];
                xPos = xPos - 2.0f * f_spacing(xPos);
              }
            }
            if (yPos - I->yPosition[yIndex] <= 0.0f && dc[1] > 0.0f) {
              yPos = yPos - f_spacing(yPos);
              yIndex = yIndex - 1;
              if (yIndex <= 0) {
                yIndex = I->ny;
                yPos = I->xPosition[yIndex <= I->nx + 1 ? yIndex : I->nx + 1];
                yPos = xPos - 2.0f * f_spacing(yPos);
              }
            }
            if (zPos - I->zPosition[zIndex] <= 0.0f && dc[2] > 0.0f) {
              zPos = zPos - f_spacing(zPos);
              zIndex = zIndex - 1;
            }
          }
          cumTmp[1] = 0.0f;
          for (int c = 1; c <= I->nc; c++) cumTmp[c + 1] = F4(I, cumulativeExt, xIndex, yIndex, zIndex, c);
          component = findIndex1(orc_mt_real(rng), cumTmp, I->nc + 1, 0);
          I->cnt.rng_draws++;
          if (component < 1) component = 1;
          if (component > I->nc) component = I->nc;
          I->cnt.collisions++;
          ssa = F4(I, ssa, xIndex, yIndex, zIndex, component);
          if (ssa < 1.0f) {
            F2(I, fluxAbsorbed, xIndex, yIndex) += photonWeight * (1.0f - ssa);
            I->volumeAbsorption[((size_t)(zIndex - 1) * I->ny + (yIndex - 1)) * I->nx + (xIndex - 1)] +=
                photonWeight * (1.0f - ssa);
            photonWeight = photonWeight * ssa;
            I->cnt.absorptions++;
          }
          if (I->computeIntensity) {
            computeIntensityContribution(I, photonWeight, xPos, yPos, zPos, xIndex, yIndex, zIndex, dc, component,
                                         rng, scatteringOrder, contributions, xIndexF, yIndexF);
            for (int i = 0; i < nD; i++) {
              size_t col = (size_t)(yIndexF[i] - 1) * I->nx + (xIndexF[i] - 1);
              I->intensity[(size_t)i * ncol + col] += contributions[i];
              I->intensityByComponent[((size_t)component * nD + i) * ncol + col] += contributions[i];
              if (contributions[i] != 0.0f) I->cnt.contributions++;
            }
          }
          if (I->useRussianRoulette && photonWeight < I->RussianRouletteW / 2.0f) {
            I->cnt.rng_draws++;
            if (orc_mt_real(rng) >= photonWeight / I->RussianRouletteW) {
              photonWeight = 0.0f;
              I->cnt.roulette_kills++;
            } else {
              photonWeight = I->RussianRouletteW;
            }
          }
          if (photonWeight <= F_TINY) goto nextPhoton;
          phaseFunctionIndex = F4(I, phaseFunctionIndex, xIndex, yIndex, zIndex, component);
          {
            const matrix* M = &I->inversePhaseFunctions[component - 1];
            scatteringAngle = computeScatteringAngle(orc_mt_real(rng), M->values + (size_t)(phaseFunctionIndex - 1) * M->numX, M->numX);
            I->cnt.rng_draws++;
          }
          next_direct(rng, &I->cnt.rng_draws, cosf(scatteringAngle), dc);
        }
      }
      continue;
    }
  nextPhoton:;
  }
  free(cumTmp);
  I->cnt.photons += nPhotons;
  I->cnt.bad += nBad;
  if (nPhotons > 0) {
    *numPhotonsProcessed = nPhotons;
    set_msg(I, "computeRadiativeTransfer: finished with photons");
    return I3RC_SUCCESS;
  }
  set_msg(I, "computeRadiativeTransfer: Didn't process any photons.");
  return I3RC_FAILURE;
}

/* MCRT:262-398 computeRadiativeTransfer.  The caller's RNG is seeded from the seed vector, used first
 * by the photon source and then by the integrator, exactly like DRV:277-287. */
static int computeRadiativeTransfer_rng(orc_integrator* I, uint32_t* rng, stream* ph) {
  if (!I->readyToCompute) {
    set_msg(I, "computeRadiativeTransfer: problem not completely specified.");
    return I3RC_FAILURE;
  }
  int numX = I->nx, numY = I->ny, numZ = I->nz, numComponents = I->nc;
  size_t ncol = (size_t)numX * numY, ncell = ncol * numZ;
  int nD = I->computeIntensity ? I->nDir : 0;
  memset(I->fluxUp, 0, sizeof(float) * ncol);
  memset(I->fluxDown, 0, sizeof(float) * ncol);
  memset(I->fluxAbsorbed, 0, sizeof(float) * ncol);
  memset(I->volumeAbsorption, 0, sizeof(float) * ncell);
  if (I->intensity) memset(I->intensity, 0, sizeof(float) * ncol * I->nDir);
  if (I->intensityByComponent) memset(I->intensityByComponent, 0, sizeof(float) * ncol * I->nDir * (numComponents + 1));
  if (I->intensityExcess) memset(I->intensityExcess, 0, sizeof(float) * (size_t)(numComponents + 1) * I->nDir);

  int rc = tabulateInversePhaseFunctions(I);
  if (rc != I3RC_FAILURE && I->computeIntensity) rc = tabulateForwardPhaseFunctions(I);
  if (rc == I3RC_FAILURE) return rc;
  int64_t numPhotonsProcessed = 0;
  rc = computeRT(I, rng, ph, &numPhotonsProcessed);
  if (rc == I3RC_FAILURE) return rc;

  if (I->computeIntensity && I->limitIntensityContributions) { /* MCRT:327-347 */
    for (int j = 0; j <= numComponents; j++)
      for (int d = 0; d < nD; d++) {
        float ex = I->intensityExcess[(size_t)j * nD + d];
        if (ex > 0.0f) {
          float* byc = I->intensityByComponent + ((size_t)j * nD + d) * ncol;
          float* in = I->intensity + (size_t)d * ncol;
          float s = 0.0f;
          for (size_t k = 0; k < ncol; k++) s += byc[k];
          for (size_t k = 0; k < ncol; k++) in[k] = in[k] + (byc[k] / s) * ex;
          for (size_t k = 0; k < ncol; k++) byc[k] = byc[k] + (byc[k] / s) * ex;
        }
      }
  }
  float* nPPC = (float*)malloc(sizeof(float) * ncol);
  if (I->xyRegularlySpaced) {
    for (size_t k = 0; k < ncol; k++) nPPC[k] = (float)numPhotonsProcessed / (float)(numX * numY);
  } else { /* MCRT:358-366 */
    for (int j = 1; j <= numY; j++)
      for (int i = 1; i <= numX; i++) {
        float v = ((I->yPosition[j + 1] - I->yPosition[j]) * (I->xPosition[i + 1] - I->xPosition[i])) /
                  ((I->xPosition[numX + 1] - I->xPosition[1]) * (I->yPosition[numY + 1] - I->yPosition[1]));
        nPPC[(size_t)(j - 1) * numX + (i - 1)] = v * (float)numPhotonsProcessed;
      }
  }
  for (size_t k = 0; k < ncol; k++) {
    I->fluxUp[k] = I->fluxUp[k] / nPPC[k];
    I->fluxDown[k] = I->fluxDown[k] / nPPC[k];
    I->fluxAbsorbed[k] = I->fluxAbsorbed[k] / nPPC[k];
  }
  for (int k = 1; k <= numZ; k++) {
    float dz = I->zPosition[k + 1] - I->zPosition[k];
    float* va = I->volumeAbsorption + (size_t)(k - 1) * ncol;
    for (size_t c = 0; c < ncol; c++) va[c] = va[c] / (nPPC[c] * dz);
  }
  if (I->computeIntensity) { /* MCRT:386-395: component 0 (surface) is not normalised (forall j = 1:numComponents) */
    for (int d = 0; d < nD; d++)
      for (size_t c = 0; c < ncol; c++) I->intensity[(size_t)d * ncol + c] = I->intensity[(size_t)d * ncol + c] / nPPC[c];
    for (int j = 1; j <= numComponents; j++)
      for (int d = 0; d < nD; d++) {
        float* byc = I->intensityByComponent + ((size_t)j * nD + d) * ncol;
        for (size_t c = 0; c < ncol; c++) byc[c] = byc[c] / nPPC[c];
      }
  }
  free(nPPC);
  return I3RC_SUCCESS;
}

int orc_computeRadiativeTransfer(orc_integrator* I, const orc_photon_source* src, const int32_t* seed, int nseed) {
  uint32_t rng[MT_N + 1];
  if (nseed == 1)
    orc_mt_seed_scalar(rng, seed[0]);
  else
    orc_mt_seed_vector(rng, seed, nseed);
  stream ph;
  memset(&ph, 0, sizeof(ph));
  int rc = new_PhotonStream(src, rng, &ph, I);
  if (rc == I3RC_FAILURE) return rc;
  rc = computeRadiativeTransfer_rng(I, rng, &ph);
  finalize_PhotonStream(&ph);
  return rc;
}

/* MCRT:711-826 reportResults (NULL = argument not present) */
int orc_reportResults(orc_integrator* I, float* meanFluxUp, float* meanFluxDown, float* meanFluxAbsorbed,
                      float* fluxUp, float* fluxDown, float* fluxAbsorbed, float* absorbedProfile,
                      float* volumeAbsorption, float* meanIntensity, float* intensity) {
  size_t ncol = (size_t)I->nx * I->ny, ncell = ncol * I->nz;
  int numColumns = (int)ncol;
  float s;
  if (meanFluxUp) {
    s = 0.0f;
    for (size_t k = 0; k < ncol; k++) s += I->fluxUp[k];
    *meanFluxUp = s / (float)numColumns;
  }
  if (meanFluxDown) {
    s = 0.0f;
    for (size_t k = 0; k < ncol; k++) s += I->fluxDown[k];
    *meanFluxDown = s / (float)numColumns;
  }
  if (meanFluxAbsorbed) {
    s = 0.0f;
    for (size_t k = 0; k < ncol; k++) s += I->fluxAbsorbed[k];
    *meanFluxAbsorbed = s / (float)numColumns;
  }
  if (fluxUp) memcpy(fluxUp, I->fluxUp, sizeof(float) * ncol);
  if (fluxDown) memcpy(fluxDown, I->fluxDown, sizeof(float) * ncol);
  if (fluxAbsorbed) memcpy(fluxAbsorbed, I->fluxAbsorbed, sizeof(float) * ncol);
  if (absorbedProfile)
    for (int k = 0; k < I->nz; k++) {
      s = 0.0f;
      for (size_t c = 0; c < ncol; c++) s += I->volumeAbsorption[(size_t)k * ncol + c];
      absorbedProfile[k] = s / (float)numColumns;
    }
  if (volumeAbsorption) memcpy(volumeAbsorption, I->volumeAbsorption, sizeof(float) * ncell);
  if (meanIntensity) {
    if (!I->intensity) {
      set_msg(I, "reportResults: intensity information not available");
      return I3RC_FAILURE;
    }
    for (int d = 0; d < I->nDir; d++) {
      s = 0.0f;
      for (size_t c = 0; c < ncol; c++) s += I->intensity[(size_t)d * ncol + c];
      meanIntensity[d] = s / (float)numColumns;
    }
  }
  if (intensity) {
    if (!I->intensity) {
      set_msg(I, "reportResults: intensity information not available");
      return I3RC_FAILURE;
    }
    memcpy(intensity, I->intensity, sizeof(float) * ncol * I->nDir);
  }
  return I3RC_SUCCESS;
}
int orc_get_intensityByComponent(orc_integrator* I, float* out) {
  if (!I->intensityByComponent) return I3RC_FAILURE;
  memcpy(out, I->intensityByComponent, sizeof(float) * (size_t)I->nx * I->ny * I->nDir * (I->nc + 1));
  return I3RC_SUCCESS;
}

/* ------------------------------------------------------------------------------------------
 * Deterministic sub-path probes
 * ---------------------------------------------------------------------------------------- */
int orc_trace_rays(orc_integrator* I, int n, const float* pos, const float* dir, const float* tauLimit, float* tauOut,
                   float* posOut, int32_t* idxOut) {
  int64_t dummy = 0;
  for (int r = 0; r < n; r++) {
    float x = pos[3 * r], y = pos[3 * r + 1], z = pos[3 * r + 2];
    int xi = 1, yi = 1, zi = 1;
    findXYIndicies(I, x, y, &xi, &yi);
    findZIndex(I, z, &zi);
    float tau;
    accumulateExtinctionAlongPath(I, dir + 3 * r, &x, &y, &z, &xi, &yi, &zi, &tau, tauLimit != NULL,
                                  tauLimit ? tauLimit[r] : 0.0f, &dummy);
    tauOut[r] = tau;
    if (posOut) {
      posOut[3 * r] = x;
      posOut[3 * r + 1] = y;
      posOut[3 * r + 2] = z;
    }
    if (idxOut) {
      idxOut[3 * r] = xi;
      idxOut[3 * r + 1] = yi;
      idxOut[3 * r + 2] = zi;
    }
  }
  return I3RC_SUCCESS;
}
int orc_sample_scattering_angles(orc_integrator* I, int comp, int entry, int n, const float* xi, float* theta) {
  if (!I->inversePhaseFunctions || !I->inversePhaseFunctions[comp].values) return I3RC_FAILURE;
  const matrix* M = &I->inversePhaseFunctions[comp];
  for (int i = 0; i < n; i++) theta[i] = computeScatteringAngle(xi[i], M->values + (size_t)entry * M->numX, M->numX);
  return I3RC_SUCCESS;
}
int orc_lookup_phase_function(orc_integrator* I, int comp, int entry, int which, int n, const float* angles, float* out) {
  matrix* arr = which == 1 ? I->tabulatedPhaseFunctions : I->tabulatedOrigPhaseFunctions;
  if (!arr || !arr[comp].values) return I3RC_FAILURE;
  const matrix* M = &arr[comp];
  for (int i = 0; i < n; i++) out[i] = lookUpPhaseFuncValFromTable(M->values + (size_t)entry * M->numX, M->numX, angles[i]);
  return I3RC_SUCCESS;
}

/* ------------------------------------------------------------------------------------------
 * DRV:264-378 batch loop and moments.  The reference accumulates the moments in default REAL; the
 * oracle accumulates them in double so that it can serve as the judge of the product's statistics
 * (documented deviation; the per-batch values themselves are the float32 ones).
 * OpenMP threads stand in for MPI ranks (one private integrator copy per thread).
 * ---------------------------------------------------------------------------------------- */
static void add_counters(orc_counters* a, const orc_counters* b) {
  int64_t* pa = (int64_t*)a;
  const int64_t* pb = (const int64_t*)b;
  for (size_t i = 0; i < sizeof(orc_counters) / sizeof(int64_t); i++) pa[i] += pb[i];
}
int orc_run_batches(const orc_integrator* proto, const orc_photon_source* src, int32_t iseed, int seedOrder,
                    int batchBegin, int nBatches, int nThreads, orc_batch_stats* st, orc_counters* counters) {
  int nx = proto->nx, ny = proto->ny, nz = proto->nz;
  int nD = proto->computeIntensity ? proto->nDir : 0;
  size_t ncol = (size_t)nx * ny, ncell = ncol * nz;
  int failed = 0;
  orc_counters total;
  memset(&total, 0, sizeof(total));
#ifdef _OPENMP
  if (nThreads <= 0) nThreads = omp_get_max_threads();
#else
  nThreads = 1;
#endif
  if (nThreads > nBatches) nThreads = nBatches;
#pragma omp parallel num_threads(nThreads)
  {
    orc_integrator* I = NULL;
    orc_copy_Integrator(proto, &I);
    memset(&I->cnt, 0, sizeof(I->cnt));
    float* fu = (float*)malloc(sizeof(float) * ncol);
    float* fd = (float*)malloc(sizeof(float) * ncol);
    float* fa = (float*)malloc(sizeof(float) * ncol);
    float* ap = (float*)malloc(sizeof(float) * nz);
    float* av = st->absorbedVolume ? (float*)malloc(sizeof(float) * ncell) : NULL;
    float* rad = nD ? (float*)malloc(sizeof(float) * ncol * nD) : NULL;
    float* mrad = nD ? (float*)malloc(sizeof(float) * nD) : NULL;
#pragma omp for schedule(dynamic, 1)
    for (int b = 0; b < nBatches; b++) {
      int batch = batchBegin + b;
      int32_t seed[2];
      if (seedOrder == 0) {
        seed[0] = iseed;
        seed[1] = batch;
      } else {
        seed[0] = batch;
        seed[1] = iseed;
      }
      int rc = orc_computeRadiativeTransfer(I, src, seed, 2);
      if (rc == I3RC_FAILURE) {
#pragma omp atomic write
        failed = 1;
        continue;
      }
      float mu_, md_, ma_;
      orc_reportResults(I, &mu_, &md_, &ma_, fu, fd, fa, ap, av, mrad, rad);
#pragma omp critical
      {
        st->meanFluxUp[0] += mu_;
        st->meanFluxUp[1] += (double)mu_ * mu_;
        st->meanFluxDown[0] += md_;
        st->meanFluxDown[1] += (double)md_ * md_;
        st->meanFluxAbsorbed[0] += ma_;
        st->meanFluxAbsorbed[1] += (double)ma_ * ma_;
        for (size_t k = 0; k < ncol; k++) {
          st->fluxUp[k] += fu[k];
          st->fluxUp[ncol + k] += (double)fu[k] * fu[k];
          st->fluxDown[k] += fd[k];
          st->fluxDown[ncol + k] += (double)fd[k] * fd[k];
          st->fluxAbsorbed[k] += fa[k];
          st->fluxAbsorbed[ncol + k] += (double)fa[k] * fa[k];
        }
        for (int k = 0; k < nz; k++) {
          st->absorbedProfile[k] += ap[k];
          st->absorbedProfile[nz + k] += (double)ap[k] * ap[k];
        }
        if (av)
          for (size_t k = 0; k < ncell; k++) {
            st->absorbedVolume[k] += av[k];
            st->absorbedVolume[ncell + k] += (double)av[k] * av[k];
          }
        if (nD && st->radiance)
          for (size_t k = 0; k < ncol * nD; k++) {
            st->radiance[k] += rad[k];
            st->radiance[ncol * nD + k] += (double)rad[k] * rad[k];
          }
        if (nD && st->meanRadiance)
          for (int d = 0; d < nD; d++) {
            st->meanRadiance[d] += mrad[d];
            st->meanRadiance[nD + d] += (double)mrad[d] * mrad[d];
          }
      }
    }
#pragma omp critical
    add_counters(&total, &I->cnt);
    free(fu);
    free(fd);
    free(fa);
    free(ap);
    free(av);
    free(rad);
    free(mrad);
    orc_finalize_Integrator(I);
  }
  if (counters) *counters = total;
  return failed ? I3RC_FAILURE : I3RC_SUCCESS;
}
