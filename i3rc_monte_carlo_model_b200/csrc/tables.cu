// Phase-function table builders on the device (setup path, run once per integrator):
//   computeInversePhaseFuncTable / computeInversePhaseFunction   Code/inversePhaseFunctions.f95:28-176
//   computeLobattoTerms, computeLegendrePolynomials, findIndex    Code/numericUtilities.f95:15-102,175-248
//   getPhaseFunctionValues_one / _table, normalizePhaseFunction   Code/scatteringPhaseFunctions.f95:446-648,1329-1345
//   tabulateForwardPhaseFunctions, computeHydridPhaseFunctions    Integrators/monteCarloRadiativeTransfer.f95:1863-2039
//
// This translation unit is compiled with -fmad=false: the deterministic sub-paths must agree with the
// reference's float32 arithmetic to 1e-5 relative, and the Legendre sums cancel to ~1e-2 of their largest
// term in the back-scattering directions, so operation order and rounding are kept (no FMA contraction).
// Parallelism: one thread per table element for the O(nSteps * nMoments) parts; the short serial parts
// (CDF running sum, hybrid transition search, normalisation) run as one thread per table entry.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tables.cuh"

namespace i3rc {

static const float Pi = 3.14159265358979312f;
#define T_TINY 1.17549435e-38f
#define T_HUGE 3.402823466e+38f

__device__ __forceinline__ float f_spacing(float x) {
  if (x == 0.0f) return T_TINY;
  float s = ldexpf(1.0f, ilogbf(fabsf(x)) - 23);
  return s < T_TINY ? T_TINY : s;
}

// largest k in [1, n] (1-based) with table[k] <= v; 0 if v < table[1]  (numericUtilities.f95:195-248; the
// hunt phase of the reference only speeds the search up, the result is the same for a monotone table)
__device__ int find_index(float v, const float* table1, int n) {
  int lo = 0, hi = n;
  for (;;) {
    if (lo == n || hi <= lo + 1) break;
    int mid = (lo + hi) / 2;
    if (v >= table1[mid])
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// P_{L}(mu) and P_{L-1}(mu) by the reference's recursion (numericUtilities.f95:185-192)
__device__ void legendre_pair(int L, float mu, float* pL, float* pLm1) {
  float p0 = 1.0f, p1 = mu;
  if (L == 0) {
    *pL = 1.0f;
    *pLm1 = 0.0f;
    return;
  }
  for (int l = 1; l <= L - 1; l++) {
    float p2 = (((float)(2 * l + 1) * mu) * p1 - (float)l * p0) / (float)(l + 1);
    p0 = p1;
    p1 = p2;
  }
  *pL = p1;
  *pLm1 = p0;
}

// ---- Lobatto abscissas (numericUtilities.f95:15-102): one block per entry, one thread per root ----
__global__ void k_lobatto(const int* __restrict__ offsets, float* __restrict__ mus, int stride) {
  int e = blockIdx.x;
  int nMom = offsets[e + 1] - offsets[e];
  int nTerms = nMom > 2 ? nMom : 2;
  float* m = mus + (size_t)e * stride - 1;  // 1-based
  int midPoint = (nTerms + 1) / 2;
  int nr = midPoint - 1;
  float pi = acosf(-1.0f);
  float c1 = (nTerms % 2 == 1) ? 1.0f : 0.5f;
  for (int i = 1 + threadIdx.x; i <= nr; i += blockDim.x) {
    float t = sinf(pi * ((float)i - c1) / ((float)nTerms - 1.0f + 0.5f));
    float last = t;
    for (int it = 0; it <= 26; it++) {
      if (it > 0 && fabsf(t - last) <= 3.0f * f_spacing(t)) break;
      float pn1, pn2;
      legendre_pair(nTerms - 1, t, &pn1, &pn2);
      float deriv = (float)(nTerms - 1) * (t * pn1 - pn2) / (t * t - 1.0f);
      float second = (2.0f * t * deriv - ((float)(nTerms * (nTerms - 1)) * pn1)) / (1.0f - t * t);
      last = t;
      t = t - deriv / second;
    }
    m[midPoint - i + 1] = -t;
  }
  if (threadIdx.x == 0) m[1] = -1.0f;
  __syncthreads();
  // symmetric half (numericUtilities.f95:91-98)
  if (nTerms % 2 == 0) {
    for (int i = 1 + threadIdx.x; i <= midPoint; i += blockDim.x) m[midPoint + i] = -m[midPoint - i + 1];
  } else {
    // reads the lower half (final) and writes the upper half; i = 0 negates the middle node in place
    for (int i = threadIdx.x; i <= midPoint - 1; i += blockDim.x) m[midPoint + i] = -m[midPoint - i];
  }
}

// value of one Legendre-series phase function at mu (getPhaseFunctionValues_one flavour:
// sum_l (c_l*(2l+1)) * P_l, scatteringPhaseFunctions.f95:483-496)
__device__ float legendre_value_one(const float* coefs, int maxL, float mu) {
  if (maxL == 0) return 1.0f / 2.0f;  // quirk Q5
  float p0 = 1.0f, p1 = mu;
  float s = 0.0f;
  s += (1.0f * 1.0f) * p0;
  s += (coefs[0] * 3.0f) * p1;
  for (int l = 1; l <= maxL - 1; l++) {
    float p2 = (((float)(2 * l + 1) * mu) * p1 - (float)l * p0) / (float)(l + 1);
    s += (coefs[l] * (float)(2 * (l + 1) + 1)) * p2;
    p0 = p1;
    p1 = p2;
  }
  return s;
}
// getPhaseFunctionValues_table flavour: sum_l c_l * ((2l+1)*P_l)  (scatteringPhaseFunctions.f95:580-584, 625-626)
__device__ float legendre_value_table(const float* coefs, int maxL, float mu) {
  if (maxL == 0) return 1.0f / 2.0f;
  float p0 = 1.0f, p1 = mu;
  float s = 0.0f;
  s += 1.0f * (1.0f * p0);
  s += coefs[0] * (3.0f * p1);
  for (int l = 1; l <= maxL - 1; l++) {
    float p2 = (((float)(2 * l + 1) * mu) * p1 - (float)l * p0) / (float)(l + 1);
    s += coefs[l] * ((float)(2 * (l + 1) + 1) * p2);
    p0 = p1;
    p1 = p2;
  }
  return s;
}

// values(i) at the Lobatto nodes (inversePhaseFunctions.f95:113-114)
__global__ void k_values_at_nodes(const int* __restrict__ offsets, const float* __restrict__ coefs,
                                  const float* __restrict__ mus, float* __restrict__ values, int stride) {
  int e = blockIdx.y;
  int nMom = offsets[e + 1] - offsets[e];
  int nA = nMom > 2 ? nMom : 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nA) return;
  float mu = cosf(acosf(mus[(size_t)e * stride + i]));
  values[(size_t)e * stride + i] = legendre_value_one(coefs + offsets[e], nMom, mu);
}

// tabulated entries: mus(i) = cos(angle(n+1-i)), values reversed (inversePhaseFunctions.f95:96-100)
__global__ void k_nodes_from_tabulated(const float* __restrict__ angles, const float* __restrict__ vals, int nA,
                                       float* __restrict__ mus, float* __restrict__ values, int stride) {
  int e = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nA) return;
  mus[(size_t)e * stride + i] = cosf(angles[nA - 1 - i]);
  values[(size_t)e * stride + i] = vals[(size_t)e * nA + (nA - 1 - i)];
}

// CDF by trapezoid in mu, normalised to [0,1] (inversePhaseFunctions.f95:122-129): one thread per entry
__global__ void k_cdf(const int* __restrict__ nAnglesPerEntry, const float* __restrict__ mus,
                      const float* __restrict__ values, float* __restrict__ cdf, int stride, int nEntries) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEntries) return;
  int nA = nAnglesPerEntry[e];
  const float* m = mus + (size_t)e * stride;
  const float* v = values + (size_t)e * stride;
  float* c = cdf + (size_t)e * stride;
  c[0] = 0.0f;
  for (int i = 1; i < nA; i++) c[i] = c[i - 1] + (m[i] - m[i - 1]) * 0.5f * (v[i] + v[i - 1]);
  float tot = c[nA - 1];
  for (int i = 0; i < nA; i++) c[i] = c[i] / tot;
}

// analytic inversion of the trapezoid CDF (inversePhaseFunctions.f95:131-170): one thread per table step
__global__ void k_invert(const int* __restrict__ nAnglesPerEntry, const float* __restrict__ mus,
                         const float* __restrict__ values, const float* __restrict__ cdf, int stride, int nSteps,
                         float* __restrict__ out) {
  int e = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // Fortran i - 1
  if (i >= nSteps) return;
  float* T = out + (size_t)e * nSteps;
  if (i == nSteps - 1) {
    T[i] = 0.0f;
    return;
  }
  int nA = nAnglesPerEntry[e];
  const float* m = mus + (size_t)e * stride - 1;  // 1-based views
  const float* v = values + (size_t)e * stride - 1;
  const float* c = cdf + (size_t)e * stride - 1;
  float p = (float)i / (float)(nSteps - 1);
  int k = find_index(p, c, nA);
  if (k < 1) k = 1;
  if (k > nA - 1) k = nA - 1;
  float r;
  if (c[k + 1] - c[k] <= f_spacing(c[k])) {
    r = acosf(m[k]);
  } else if (fabsf(v[k] - v[k + 1]) <= f_spacing(v[k])) {
    r = acosf(m[k] + (m[k + 1] - m[k]) * (p - c[k]) / (c[k + 1] - c[k]));
  } else {
    r = acosf(m[k] + (m[k + 1] - m[k]) / (v[k] - v[k + 1]) *
                         (v[k] - sqrtf(((c[k + 1] - p) * (v[k] * v[k]) + (p - c[k]) * (v[k + 1] * v[k + 1])) /
                                       (c[k + 1] - c[k]))));
  }
  T[i] = r;
}

// forward tables on nSteps equal angle steps (MCRT:1896-1903)
__global__ void k_forward_legendre(const int* __restrict__ offsets, const float* __restrict__ coefs, int nSteps,
                                   float* __restrict__ out) {
  int e = blockIdx.y;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nSteps) return;
  float ang = (float)j / (float)(nSteps - 1) * Pi;
  out[(size_t)e * nSteps + j] = legendre_value_table(coefs + offsets[e], offsets[e + 1] - offsets[e], cosf(ang));
}
__global__ void k_forward_tabulated(const float* __restrict__ angles, const float* __restrict__ vals, int nA,
                                    int nEntries, int nSteps, float* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nSteps) return;
  float ang = (float)j / (float)(nSteps - 1) * Pi;
  const float* sa = angles - 1;
  int idx = find_index(ang, sa, nA);
  if (idx < 1) idx = 1;
  int idx1 = idx + 1;
  float dMu;
  if (idx < nA) {
    dMu = cosf(sa[idx1]) - cosf(sa[idx]);
  } else {
    dMu = T_HUGE;
    idx1 = idx;
  }
  float w = 1.0f - (cosf(ang) - cosf(sa[idx])) / dMu;  // scatteringPhaseFunctions.f95:604-613
  for (int e = 0; e < nEntries; e++) {
    const float* v = vals + (size_t)e * nA - 1;
    out[(size_t)e * nSteps + j] = w * v[idx] + (1.0f - w) * v[idx1];
  }
}

// normalizePhaseFunction (scatteringPhaseFunctions.f95:1329-1345): one thread per entry, in place
__global__ void k_normalize(const float* __restrict__ angles, float* __restrict__ vals, int nA, int nEntries) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEntries) return;
  float* v = vals + (size_t)e * nA;
  float dot = 0.0f;
  for (int i = 0; i < nA - 1; i++) dot += (cosf(angles[i + 1]) - cosf(angles[i])) * (0.5f * (v[i + 1] + v[i]));
  for (int i = 0; i < nA; i++) v[i] = -v[i] * 2.0f / dot;
}

// ---- hybrid phase functions (MCRT:1925-2039): one thread per entry ----
__device__ float hyb_norm(const float* ac, const float* v, const float* g, int nA, int ti) {
  float ig = 0.0f, io = 0.0f;
  for (int k = 1; k <= ti - 1; k++) ig += (0.5f * (g[k] + g[k + 1])) * (ac[k] - ac[k + 1]);
  for (int k = ti; k <= nA - 1; k++) io += (0.5f * (v[k] + v[k + 1])) * (ac[k] - ac[k + 1]);
  if (io >= 2.0f) return 1.0f / ig;
  return (2.0f - io) / ig;
}
__device__ float hyb_diff(const float* ac, const float* v, const float* g, int nA, int ti) {
  return hyb_norm(ac, v, g, nA, ti) * g[ti] - v[ti];
}
__global__ void k_hybrid_prepare(int nA, float width, float* __restrict__ ac, float* __restrict__ gaus) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nA) return;
  float ang = (float)j / (float)(nA - 1) * Pi;
  ac[j] = cosf(ang);
  float r = ang / (width * Pi / 180.0f);
  gaus[j] = expf(-(r * r));
}
__global__ void k_hybrid(const float* __restrict__ orig, float* __restrict__ out, int nA, int nEntries, float width,
                         const float* __restrict__ ac0, const float* __restrict__ gaus0, int firstLowerBound) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEntries) return;
  const float* v = orig + (size_t)e * nA - 1;
  float* nv = out + (size_t)e * nA - 1;
  const float* ac = ac0 - 1;
  const float* g = gaus0 - 1;
  int lowerBound = firstLowerBound;
  // "exit entryLoop" of the reference leaves this and all later entries untouched; the bound does not
  // depend on the entry, so every thread takes the same decision
  if (lowerBound >= nA - 2) return;
  float lowDiff = hyb_diff(ac, v, g, nA, lowerBound);
  int increment = 1, upperBound;
  float upDiff;
  for (;;) {
    upperBound = lowerBound + increment < nA - 1 ? lowerBound + increment : nA - 1;
    upDiff = hyb_diff(ac, v, g, nA, upperBound);
    if (lowerBound == nA - 1) return;  // no root: keep the original
    if (lowDiff * upDiff < 0.0f) break;
    lowerBound = upperBound;
    lowDiff = upDiff;
    increment *= 2;
  }
  for (;;) {
    if (upperBound <= lowerBound + 1) break;
    int mid = (lowerBound + upperBound) / 2;
    float midDiff = hyb_diff(ac, v, g, nA, mid);
    if (midDiff * upDiff < 0.0f) {
      lowerBound = mid;
      lowDiff = midDiff;
    } else {
      upperBound = mid;
      upDiff = midDiff;
    }
  }
  int ti = lowerBound;
  float P0 = hyb_norm(ac, v, g, nA, ti);
  for (int k = 1; k <= ti; k++) nv[k] = P0 * g[k];
}

// ---- host launchers --------------------------------------------------------------------------------
#define CK(x)                       \
  do {                              \
    cudaError_t e_ = (x);           \
    if (e_ != cudaSuccess) return e_; \
  } while (0)

cudaError_t normalize_tabulated(const float* dAngles, float* dValues, int nAngles, int nEntries, int repeats,
                                cudaStream_t st) {
  for (int r = 0; r < repeats; r++) k_normalize<<<(nEntries + 63) / 64, 64, 0, st>>>(dAngles, dValues, nAngles, nEntries);
  return cudaGetLastError();
}

cudaError_t build_inverse_table(const PhaseTableDev& t, int nSteps, float* dOut, cudaStream_t st) {
  int stride = t.maxNodes;
  float *mus = nullptr, *values = nullptr, *cdf = nullptr;
  int* nA = nullptr;
  size_t bytes = sizeof(float) * (size_t)stride * t.nEntries;
  CK(cudaMalloc(&mus, bytes));
  CK(cudaMalloc(&values, bytes));
  CK(cudaMalloc(&cdf, bytes));
  CK(cudaMalloc(&nA, sizeof(int) * t.nEntries));
  CK(cudaMemcpyAsync(nA, t.hostNodesPerEntry, sizeof(int) * t.nEntries, cudaMemcpyHostToDevice, st));
  if (t.kind == 1) {
    k_lobatto<<<t.nEntries, 128, 0, st>>>(t.offsets, mus, stride);
    dim3 g((stride + 127) / 128, t.nEntries);
    k_values_at_nodes<<<g, 128, 0, st>>>(t.offsets, t.coefs, mus, values, stride);
  } else {
    dim3 g((t.nAngles + 127) / 128, t.nEntries);
    k_nodes_from_tabulated<<<g, 128, 0, st>>>(t.angles, t.values, t.nAngles, mus, values, stride);
  }
  k_cdf<<<(t.nEntries + 31) / 32, 32, 0, st>>>(nA, mus, values, cdf, stride, t.nEntries);
  dim3 gi((nSteps + 127) / 128, t.nEntries);
  k_invert<<<gi, 128, 0, st>>>(nA, mus, values, cdf, stride, nSteps, dOut);
  cudaError_t err = cudaGetLastError();
  cudaStreamSynchronize(st);
  cudaFree(mus);
  cudaFree(values);
  cudaFree(cdf);
  cudaFree(nA);
  return err;
}

cudaError_t build_forward_table(const PhaseTableDev& t, int nSteps, float* dOut, cudaStream_t st) {
  if (t.kind == 1) {
    dim3 g((nSteps + 127) / 128, t.nEntries);
    k_forward_legendre<<<g, 128, 0, st>>>(t.offsets, t.coefs, nSteps, dOut);
  } else {
    k_forward_tabulated<<<(nSteps + 127) / 128, 128, 0, st>>>(t.angles, t.values, t.nAngles, t.nEntries, nSteps, dOut);
  }
  return cudaGetLastError();
}

cudaError_t build_hybrid_table(const float* dOrig, float* dOut, int nSteps, int nEntries, float widthDeg,
                               cudaStream_t st) {
  float *ac = nullptr, *gaus = nullptr;
  CK(cudaMalloc(&ac, sizeof(float) * nSteps));
  CK(cudaMalloc(&gaus, sizeof(float) * nSteps));
  CK(cudaMemcpyAsync(dOut, dOrig, sizeof(float) * (size_t)nSteps * nEntries, cudaMemcpyDeviceToDevice, st));
  k_hybrid_prepare<<<(nSteps + 127) / 128, 128, 0, st>>>(nSteps, widthDeg, ac, gaus);
  // lowerBound = findIndex(width, angles) + 1 on the equal-angle grid (MCRT:1954), computed on the host in
  // the same float arithmetic as the table angles
  int lb = 0;
  {
    float wr = widthDeg * Pi / 180.0f;
    int lo = 0, hi = nSteps;
    for (;;) {
      if (lo == nSteps || hi <= lo + 1) break;
      int mid = (lo + hi) / 2;
      float a = (float)(mid - 1) / (float)(nSteps - 1) * Pi;
      if (wr >= a)
        lo = mid;
      else
        hi = mid;
    }
    lb = lo + 1;
  }
  k_hybrid<<<(nEntries + 31) / 32, 32, 0, st>>>(dOrig, dOut, nSteps, nEntries, widthDeg, ac, gaus, lb);
  cudaError_t err = cudaGetLastError();
  cudaStreamSynchronize(st);
  cudaFree(ac);
  cudaFree(gaus);
  return err;
}

}  // namespace i3rc
