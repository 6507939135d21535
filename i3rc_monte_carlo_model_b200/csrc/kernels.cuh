// __global__ kernels of the integrator (sm_100a).  The photon transport kernel is the hot path; the rest are
// the small setup / normalisation / reporting kernels around it.
#pragma once
#include <cuda_runtime.h>

#include "transport.cuh"

namespace i3rc {

// ---- K1: persistent, warp-cooperative photon transport ---------------------------------------------------------
// grid = (#SMs x resident blocks).  Every lane owns one photon slot and refills it from the device photon counter
// (one warp-aggregated atomicAdd per refill round).  A warp alternates between two phases:
//
//   EVENT phase  (uniform code over the lanes whose path segment has ended): boundary / collision handling, then a
//                warp-uniform loop over the radiance directions in which every such lane builds its local-estimate
//                ray as a 36-byte TASK and pushes it into the warp's shared-memory ring, then roulette + scattering
//                and the start of the next path segment;
//   TRACE phase  (rounds of a few DDA cell crossings for every lane, then ray bookkeeping for all lanes at once): a lane traces its own path segment; when that
//                ends it pops local-estimate tasks -- of ANY photon of the warp -- from the ring and traces those,
//                so lanes stay busy while the longest segment of the warp is still running.  The phase ends when the
//                ring is empty and at least `eventThreshold` lanes wait for their event.
//
// The physics functions are the ones of transport.cuh; the per-photon Philox streams make the result independent of
// which lane traces which ray (up to float summation order in the tallies).
constexpr int QCAP = 128;  // local-estimate tasks per warp (ring, power of two): 4.5 KB of shared memory per warp

template <int BLOCK, bool REG, bool FAST, int MINB, int STEPS>
__global__ void __launch_bounds__(BLOCK, MINB) k_transport(const ProblemT<REG, FAST> p, const int eventThreshold) {
  __shared__ LeTask s_task[BLOCK / 32][QCAP];
  __shared__ int s_head[BLOCK / 32];
  __shared__ uint32_t s_cnt[BLOCK / 32][CNT_N];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  LeTask* q = s_task[warp];
  int* headp = &s_head[warp];
  int tail = 0;  // warp-uniform; the ring holds positions [head, tail)
  if (lane == 0) *headp = 0;
  __syncwarp();

  Lane L;
  L.cnt = s_cnt[warp];
  if (lane < CNT_N) s_cnt[warp][lane] = 0;
  __syncwarp();
  L.active = 0;
  L.done = DONE_RUN;
  L.mode = MODE_PHOTON;
  L.nsteps = 0;
  bool hasRay = false;    // the ray registers hold a ray in flight (own segment if L.mode == MODE_PHOTON)
  bool pending = false;   // own segment ended, event not processed yet
  bool exhausted = false;

  // STEPS DDA crossings for every lane that has a running ray, then -- for all lanes at once, so that the
  // divergent bookkeeping code runs with as many lanes as possible -- finished rays are closed and idle lanes pop tasks
  auto trace_round = [&]() {
#pragma unroll 1
    for (int k = 0; k < STEPS; k++)
      if (hasRay && L.done == DONE_RUN) dda_step(p, L);
    if (hasRay && L.done != DONE_RUN) {
      if (L.mode == MODE_PHOTON) {
        segment_finished(p, L);
        pending = true;
        hasRay = false;
      } else if (!finish_le_ray(p, L)) {
        hasRay = false;
      }
    }
    if (!hasRay && *(volatile int*)headp < tail) {
      int h = atomicAdd(headp, 1);
      if (h < tail) {
        start_le_task(p, L, q[h & (QCAP - 1)]);
        hasRay = true;
      }
    }
  };
  auto fix_head = [&]() {  // pops may overshoot the tail
    __syncwarp();
    if (lane == 0 && *headp > tail) *headp = tail;
    __syncwarp();
  };

  for (;;) {
    // ================= EVENT phase =================
    const bool ev = pending && !hasRay;
    bool alive = false;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;  // the event's block of deviates
    if (ev) {
      pending = false;
      L.rng.next4(p.key0, p.key1, a0, a1, a2, a3);
      alive = photon_event(p, L, a0, a1) != 0;
    }
    if (p.computeIntensity) {
      if (__any_sync(full, alive)) {
        float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
        for (int d = 0; d < p.nDir; d++) {
          if ((d & 1) == 0 && alive) L.rng.next4(p.key0, p.key1, c0, c1, c2, c3);  // one block per two directions
          LeTask t;
          const bool push = alive && make_le_task(p, L, d, (d & 1) ? c2 : c0, (d & 1) ? c3 : c1, t);
          const unsigned m = __ballot_sync(full, push);
          const int n = __popc(m);
          if (n == 0) continue;
          int head = __shfl_sync(full, *(volatile int*)headp, 0);
          if (tail - head + n > QCAP) {
            // ring full: every lane (also the ones in the middle of their event) helps to drain it; lanes of this
            // event phase must be free again before they start their next segment
            for (;;) {
              trace_round();
              const bool stillQueued = __shfl_sync(full, *(volatile int*)headp, 0) < tail;
              if (!stillQueued && !__any_sync(full, ev && hasRay)) break;
            }
            fix_head();
          }
          if (push) q[(tail + __popc(m & lt)) & (QCAP - 1)] = t;
          tail += n;
          __syncwarp();
        }
      }
    }
    if (alive) {
      continue_photon(p, L, a1, a2, a3);  // roulette, scattering, start of the next own segment
      if (L.active) hasRay = true;
    }
    // refill finished slots
    {
      const bool need = !L.active && !exhausted && !hasRay && !pending;
      const unsigned m = __ballot_sync(full, need);
      if (m) {
        unsigned long long base = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(p.nextPhoton, (unsigned long long)__popc(m));
        base = __shfl_sync(full, base, leader);
        if (need) {
          const long long id = (long long)base + __popc(m & lt);
          if (id < p.src.n) {
            init_photon(p, L, id);
            hasRay = true;
          } else {
            exhausted = true;
          }
        }
      }
    }
    __syncwarp();
    // ================= TRACE phase =================
    bool anything = false;
    // towards the end of a batch fewer lanes hold photons: do not wait for more events than can come
    const int nActive = __popc(__ballot_sync(full, L.active != 0));
    const int threshold = min(eventThreshold, max(1, nActive >> 1));
    for (;;) {
      trace_round();
      const unsigned busy = __ballot_sync(full, hasRay);
      const int ready = __popc(__ballot_sync(full, pending && !hasRay));
      if (busy == 0) {
        anything = ready > 0;
        break;
      }
      const bool queued = __shfl_sync(full, *(volatile int*)headp, 0) < tail;
      if (!queued && ready >= threshold) {
        anything = true;
        break;
      }
    }
    fix_head();
    if (!anything && !__any_sync(full, L.active || !exhausted)) break;
  }
  // flush the warp's counters
  __syncwarp();
  if (lane < CNT_N && s_cnt[warp][lane]) atomicAdd(p.counters + lane, (unsigned long long)s_cnt[warp][lane]);
}

// ---- probes: accumulateExtinctionAlongPath for explicit rays (deterministic sub-path parity) ---------------
__global__ void k_trace_rays(const ProblemT<false> p, int n, const float* __restrict__ pos, const float* __restrict__ dir,
                             const float* __restrict__ tauLimit, float* __restrict__ tauOut,
                             float* __restrict__ posOut, int* __restrict__ idxOut) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  Lane L;
  uint32_t cnt[CNT_N];
  for (int i = 0; i < CNT_N; i++) cnt[i] = 0;
  L.cnt = cnt;
  locate_abs(p.xe, p.xyRegular, p.nx, p.x0, p.xmax, p.dx, pos[3 * r], 1, &L.cx, &L.fx);
  locate_abs(p.ye, p.xyRegular, p.ny, p.y0, p.ymax, p.dy, pos[3 * r + 1], 1, &L.cy, &L.fy);
  locate_abs(p.ze, p.zRegular, p.nz, p.z0, p.zmax, p.dz, pos[3 * r + 2], 0, &L.cz, &L.fz);
  float dx = dir[3 * r], dy = dir[3 * r + 1], dz = dir[3 * r + 2];
  start_ray(p, L, dx, dy, dz, inv_abs(dx), inv_abs(dy), inv_abs(dz), tauLimit ? tauLimit[r] : INFINITY);
  while (L.done == DONE_RUN) dda_step(p, L);
  tauOut[r] = L.done == DONE_BAD ? -2.0f : L.tau;
  ray_local(p, L, &L.fx, &L.fy, &L.fz);
  if (posOut) {
    posOut[3 * r] = abs_x(p, L.ix, L.fx);
    posOut[3 * r + 1] = abs_y(p, L.iy, L.fy);
    posOut[3 * r + 2] = L.done == DONE_TOP ? p.zmax : (L.done == DONE_BOTTOM ? p.z0 : abs_z(p, L.iz, L.fz));
  }
  if (idxOut) {
    idxOut[3 * r] = L.ix + 1;
    idxOut[3 * r + 1] = L.iy + 1;
    idxOut[3 * r + 2] = L.iz + 1;
  }
}
__global__ void k_sample_angles(const float* __restrict__ T, int nSteps, int n, const float* __restrict__ xi,
                                float* __restrict__ theta) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) theta[i] = scattering_angle(T, nSteps, xi[i]);
}
__global__ void k_lookup_phase(const float* __restrict__ T, int nSteps, int n, const float* __restrict__ ang,
                               float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = phase_lookup(T, nSteps, ang[i]);
}

// ---- setup: getOpticalPropertiesByComponent on the device (Code/opticalProperties.f95:429-539) -------------
struct ComponentDev {
  const float* ext;
  const float* ssa;
  const int* pfi;
  int uniform, zBase, nz;  // zBase 0-based
};
__global__ void k_expand_components(int nx, int ny, int nz, int nc, const ComponentDev* __restrict__ comps,
                                    float* __restrict__ totalExt, float* __restrict__ cumExt,
                                    float* __restrict__ ssa, int* __restrict__ pfIdx) {
  size_t ncell = (size_t)nx * ny * nz;
  size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  int iz = (int)(cell / ((size_t)nx * ny));
  size_t col = cell - (size_t)iz * nx * ny;
  float run = 0.0f;
  for (int c = 0; c < nc; c++) {
    ComponentDev k = comps[c];
    float e = 0.0f, s = 0.0f;
    int f = 0;
    int kz = iz - k.zBase;
    if (kz >= 0 && kz < k.nz) {
      size_t src = k.uniform ? (size_t)kz : (size_t)kz * nx * ny + col;
      e = k.ext[src];
      s = k.ssa[src];
      f = k.pfi[src];
    }
    run = (c == 0) ? e : e + run;  // cumulativeExt(:,:,:,i) + cumulativeExt(:,:,:,i-1)
    cumExt[(size_t)c * ncell + cell] = run;
    ssa[(size_t)c * ncell + cell] = s;
    pfIdx[(size_t)c * ncell + cell] = f;
  }
  totalExt[cell] = run;
  if (run > F_TINY)
    for (int c = 0; c < nc; c++) cumExt[(size_t)c * ncell + cell] = cumExt[(size_t)c * ncell + cell] / run;
}
// MCRT:233-234: nudge the last cumulative fraction to 1 + epsilon; also the domain maximum of totalExt
__global__ void k_bump_and_max(size_t ncell, float* __restrict__ lastCum, const float* __restrict__ totalExt,
                               unsigned int* __restrict__ maxBits) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float m = 0.0f;
  if (i < ncell) {
    const float eps = 1.1920929e-7f;
    float c = lastCum[i];
    if (fabsf(c - 1.0f) <= eps) lastCum[i] = 1.0f + eps;
    m = fmaxf(totalExt[i], 0.0f);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(maxBits, __float_as_uint(m));
}

// ---- post-processing of one batch (MCRT:327-395) -------------------------------------------------------------
// sum over the columns of `rows` consecutive (nx*ny)-slabs: out[r] = sum(in[r*ncol : (r+1)*ncol]) in double
__global__ void k_slab_sums(const float* __restrict__ in, size_t ncol, double* __restrict__ out) {
  __shared__ double sh[256];
  const float* row = in + (size_t)blockIdx.x * ncol;
  double s = 0.0;
  for (size_t i = threadIdx.x; i < ncol; i += blockDim.x) s += (double)row[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}
// redistribute the clipped local-estimate excess in proportion to intensityByComponent (MCRT:327-347)
__global__ void k_redistribute_excess(int nDir, int ncomp1, size_t ncol, const float* __restrict__ excess,
                                      const double* __restrict__ slabSum, float* __restrict__ intensity,
                                      float* __restrict__ intByComp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int d = blockIdx.y;
  if (i >= ncol) return;
  for (int j = 0; j < ncomp1; j++) {
    float ex = excess[j * nDir + d];
    if (ex > 0.0f) {
      size_t o = ((size_t)j * nDir + d) * ncol + i;
      float share = (intByComp[o] / (float)slabSum[j * nDir + d]) * ex;
      intensity[(size_t)d * ncol + i] += share;
      intByComp[o] += share;
    }
  }
}
struct NormArgs {
  int nx, ny, nz, nDir, nc, xyRegular, trackByComponent;
  float numPhotons;
  const float *xe, *ye, *ze;
  float *fluxUp, *fluxDown, *fluxAbs, *volAbs, *intensity, *intByComp;
};
// divide by the mean number of photons per column, absorption also by the layer depth (MCRT:353-395)
__global__ void k_normalize(const NormArgs a) {
  size_t ncol = (size_t)a.nx * a.ny;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  float nppc;
  if (a.xyRegular) {
    nppc = a.numPhotons / (float)(a.nx * a.ny);
  } else {
    int ix = (int)(i % a.nx), iy = (int)(i / a.nx);
    nppc = ((a.ye[iy + 1] - a.ye[iy]) * (a.xe[ix + 1] - a.xe[ix])) /
           ((a.xe[a.nx] - a.xe[0]) * (a.ye[a.ny] - a.ye[0]));
    nppc = nppc * a.numPhotons;
  }
  a.fluxUp[i] = a.fluxUp[i] / nppc;
  a.fluxDown[i] = a.fluxDown[i] / nppc;
  a.fluxAbs[i] = a.fluxAbs[i] / nppc;
  for (int k = 0; k < a.nz; k++) {
    float dz = a.ze[k + 1] - a.ze[k];
    a.volAbs[(size_t)k * ncol + i] = a.volAbs[(size_t)k * ncol + i] / (nppc * dz);
  }
  for (int d = 0; d < a.nDir; d++) a.intensity[(size_t)d * ncol + i] = a.intensity[(size_t)d * ncol + i] / nppc;
  if (a.trackByComponent)  // components 1..nc only, like the reference's forall (MCRT:390-394)
    for (int j = 1; j <= a.nc; j++)
      for (int d = 0; d < a.nDir; d++) {
        size_t o = ((size_t)j * a.nDir + d) * ncol + i;
        a.intByComp[o] = a.intByComp[o] / nppc;
      }
}

// ---- batch moments (monteCarloDriver.f95:300-321): s1 += x, s2 += x*x ------------------------------------------
__global__ void k_moments_f(const float* __restrict__ x, size_t n, double* __restrict__ s1, double* __restrict__ s2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double v = (double)x[i];
    s1[i] += v;
    s2[i] += v * v;
  }
}
// the same for domain means: x = float(sum / ncol) like reportResults (MCRT:739-742, 780, 803-805)
__global__ void k_moments_mean(const double* __restrict__ sums, int n, double invCols, double* __restrict__ s1,
                               double* __restrict__ s2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double v = (double)(float)(sums[i] * invCols);
    s1[i] += v;
    s2[i] += v * v;
  }
}
__global__ void k_means_to_float(const double* __restrict__ sums, int n, double invCols, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)(sums[i] * invCols);
}
// mean and standard error from the two moments (monteCarloDriver.f95:358-378)
__global__ void k_stats_finish(const double* __restrict__ s1, const double* __restrict__ s2, size_t n, double solarFlux,
                               int numBatches, double* __restrict__ mean, double* __restrict__ err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double m = solarFlux * s1[i] / numBatches;
    double q = solarFlux * s2[i] / numBatches;
    mean[i] = m;
    double v = q - m * m;
    err[i] = sqrt((v > 0.0 ? v : 0.0) / (double)(numBatches - 1));
  }
}

}  // namespace i3rc
