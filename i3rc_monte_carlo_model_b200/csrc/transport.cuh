// Photon transport core of the B200 integrator: the device-side restatement of
//   computeRT                      Integrators/monteCarloRadiativeTransfer.f95:400-707   (MCRT)
//   accumulateExtinctionAlongPath  MCRT:1654-1807
//   computeIntensityContribution   MCRT:1419-1611
//   computeScatteringAngle / lookUpPhaseFuncValsFromTable / next_direct   MCRT:1390, 1613, 2086
//   new_PhotonStream_* sampling    Code/monteCarloIllumination.f95:62-424
//   computeSurfaceReflectance      Code/surfaceProperties.f95:121-162
//
// Design (not a translation):
//  * one persistent thread per photon slot; a finished slot refills from a device photon counter;
//  * every lane is a small state machine whose only inner operation is ONE cell crossing of a 3-D DDA,
//    shared by photon path segments and local-estimate rays, so a warp's lanes run the same code even
//    when they are in different phases of a photon's life;
//  * positions are (cell index, fractional offset inside the cell); the ray is advanced in its own
//    remaining path length to the next cell face on each axis (Amanatides-Woo in cell-local form), so periodic
//    wrap-around touches indices only and no spacing()-style nudges are needed;
//  * random numbers come from a per-photon Philox4x32-10 stream (philox.cuh).
//
// Everything here is __host__ __device__ so that tests/hostsim can run the very same code on the CPU
// (one lane at a time) against the oracle before any GPU time is spent.  The product never runs it on
// the CPU: the only entry point the library exposes launches the __global__ wrappers in api.cu.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "philox.cuh"

#ifdef __CUDA_ARCH__
#define I3RC_LDG(p) __ldg(p)
#define I3RC_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define I3RC_COUNT(L, i, v) atomicAdd((L).cnt + (i), (uint32_t)(v))
#define I3RC_LOG(x) __logf(x)
#define I3RC_EXP(x) __expf(x)
#define I3RC_COS(x) __cosf(x)
#define I3RC_SINCOS(x, s, c) __sincosf((x), (s), (c))
#define I3RC_FDIV(a, b) __fdividef((a), (b))
#else
#define I3RC_LDG(p) (*(p))
#define I3RC_ATOMIC_ADD(p, v) (*(p) += (v))
#define I3RC_COUNT(L, i, v) ((L).cnt[(i)] += (uint32_t)(v))
#define I3RC_LOG(x) logf(x)
#define I3RC_EXP(x) expf(x)
#define I3RC_COS(x) cosf(x)
#define I3RC_SINCOS(x, s, c) (*(s) = sinf(x), *(c) = cosf(x))
#define I3RC_FDIV(a, b) ((a) / (b))
#endif

#ifndef __CUDACC__
struct int2 {  // (the CUDA vector types, for the CPU build of this header)
  int x, y;
};
struct float4 {
  float x, y, z, w;
};
#endif

// Build with -DI3RC_CHECK to turn the index invariants of the transport code into run-time checks: a violated check
// counts as a "bad" photon (counter CNT_BAD, which the test-suite requires to be zero) instead of touching memory.
#ifdef I3RC_CHECK
#define I3RC_ASSERT(L, cond) (!(cond) ? (void)I3RC_COUNT(L, CNT_BAD, 1) : (void)0)
#else
#define I3RC_ASSERT(L, cond) ((void)0)
#endif

namespace i3rc {

constexpr int MAX_DIRS = 32;
constexpr float F_TINY = 1.17549435e-38f;
constexpr float F_PI = 3.14159265358979312f;
constexpr int DIR_STRIDE = 8;
constexpr int MAX_RAY_STEPS = 1 << 26;  // hang protection: a ray that long is dropped as "bad"

enum { CNT_PHOTONS = 0, CNT_BAD, CNT_CROSS_PH, CNT_CROSS_LE, CNT_COLL, CNT_ABS, CNT_CONTRIB, CNT_TOP, CNT_SURF,
       CNT_RNG, CNT_KILL, CNT_NULL, CNT_SKIP, CNT_SKIP_LE,
#ifdef I3RC_SCHED_STATS  // make STATS=1: what the warps of k_transport do with their lanes (printed by fetch_counters)
       ST_ROUNDS, ST_START_RUN, ST_PAIRS, ST_LANE_PAIRS, ST_STARVED, ST_RING_LEFT, ST_BATCH, ST_BATCH_HAS, ST_BATCH_ALIVE, ST_SUSP, ST_LE_PUSH,
       ST_BIRTHS, ST_BIRTH_BATCH,
#endif
       CNT_N };
// (the value is worked out by ALL lanes -- it may be a ballot --, lane 0 adds it)
#ifdef I3RC_SCHED_STATS
#define I3RC_STAT(W, k, v)                                          \
  do {                                                              \
    const uint32_t i3rc_stat_v = (uint32_t)(v);                     \
    if ((threadIdx.x & 31) == 0) (W).cnt[k] += i3rc_stat_v;         \
  } while (0)
#else
#define I3RC_STAT(W, k, v) ((void)0)
#endif
// Empty-space codes (Problem::ext of a JUMP kernel): an empty cell whose whole Chebyshev neighbourhood of radius n + 1 is
// empty (periodic in x and y; beyond the top and the bottom counts as empty) holds -n instead of 0, for n in
// [JUMP_MIN, JUMP_MAX].  A ray that has just crossed such a cell may cross up to n further cells on every axis without
// looking at them (ray_advance_far).
constexpr int JUMP_MIN = 3, JUMP_MAX = 15;

enum { DONE_RUN = 0, DONE_INSIDE = 1, DONE_TOP = 2, DONE_BOTTOM = 3, DONE_BAD = 4,
       DONE_IDLE = 5,   // warp-cooperative kernel: the lane holds no ray
       DONE_NEW = 6,    // warp-cooperative kernel: the photon slot waits for a new photon
       DONE_STOP = 7 }; // the ray has ended; how (INSIDE / TOP / BOTTOM) is resolved by ray_after_steps()
enum { MODE_PHOTON = 0, MODE_LE_PLAIN = 1, MODE_LE_SMALL = 2, MODE_LE_BIG1 = 3, MODE_LE_BIG2 = 4 };

struct TableDesc {
  const float* inv;      // [nEntries][nInv]   scattering angle vs CDF (radians)
  const float* fwd;      // [nEntries][nFwd]   phase function on equal angle steps (hybrid if enabled)
  const float* fwdOrig;  // [nEntries][nFwd]
  int nInv, nFwd, nEntries, pad;
};

struct SourceDev {
  int kind;
  long long n;
  float mu, phi;  // directional: -|solarMu|, azimuth in radians
  float x, y, z;
  float detectorMu, detectorPhi;
  int pointsUp, hasDx, hasDy;
  float deltaX, deltaY;
  const float *ax, *ay, *az, *amu, *aphi;  // I3RC_SRC_ARRAYS (device copies)
  // I3RC_SRC_ARRAYS uploaded while the kernel runs: number of photons whose array elements have arrived so far (written
  // by the copy engine after every piece of the upload); null = everything is there before the launch
  const unsigned long long* avail;
};

struct Problem {
  int nx, ny, nz, nc;
  int xyRegular, zRegular;
  float x0, y0, z0, xmax, ymax, zmax;
  float dx, dy, dz;
  const float *xe, *ye, *ze;
  // Total extinction, the one field the rays gather from, with element strides (esx, esy, esz): cell (ix,iy,iz) is
  // ext[ix*esx + iy*esy + iz*esz].  The library keeps it z-fastest (esz = 1): most crossings of a ray are through
  // horizontal faces (layers are thinner than columns are wide, and solar / viewing directions are steep), so
  // consecutive gathers of a ray then fall into the same 32-byte sector far more often than with the x-fastest order
  // of the other fields.
  const float* ext;
  int esx, esy, esz;
  long long extN;  // number of elements of ext (checked builds)
  // Horizontally uniform layers (clear air or gas only, above and below the clouds of a typical domain) need no 3-D
  // storage: when nzc > 0 the field holds only the nzc layers that vary horizontally, ext[(ix*ny + iy)*nzc + k], and
  // zlut[iz] = { k, or -1 for a uniform layer ; bits of the uniform layer's extinction }; esz stays 1, so the cell index
  // of a ray is its column's base + iz.  A 512x512x256 field with 80 cloudy layers shrinks from 268 MB to 84 MB and
  // fits the L2 cache.
  const int2* zlut;
  int nzc;
  // Runs of consecutive horizontally uniform layers ("slabs": the clear air above and below the clouds) can be crossed
  // in one go whatever the horizontal position: zslab[iz] = { layers of the slab above iz, layers below iz, vertical
  // optical depth of those above, of those below } for a uniform layer iz (regular grids; ray_cross_slab).  Null: unused.
  const float4* zslab;
  const float* cumExt;
  const float* ssa;
  const int* pfIdx;
  float maxExt;
  const TableDesc* tables;
  int computeIntensity, nDir;
  // Radiance directions that point straight up (mu = 1: the nadir view of every I3RC case): their local-estimate rays stay
  // in the column they start in, so the optical path to the top is read from the column's suffix sums of extinction x
  // layer depth, colTau[k][iy][ix] = sum over layers >= k (k = 0 .. nz), instead of being traced (make_le_task).
  const float* colTau;
  uint32_t vertMask;  // bit d: direction d is (0, 0, 1)
  // Russian roulette for intensity: a local-estimate ray contributes only if it reaches the top within an optical-path
  // budget that is known BEFORE it is traced (MCRT:1554-1559: tauFree; :1566-1587: the first-stage limit + tauFree).
  // leLB[d][iz][iy][ix] is a LOWER BOUND of the optical path to the top along direction d from anywhere in the cell
  // (le_lower_bound below; regular grids): when it exceeds the budget the ray is known to contribute nothing and is
  // not traced.  Null: unused.  leLBBins = 8: one bound per octant of the cell the ray starts in, [d][octant][cell], octant
  // = (fx >= 1/2) + 2 (fy >= 1/2) + 4 (fz >= 1/2); 1: one bound per cell.
  const float* leLB;
  int leLBBins;
  // ... and leUB[d][cell] an UPPER bound: when even that fits the budget the ray is known to reach the top, and where its
  // contribution does not depend on the optical path (the roulette's survivors, MCRT:1556, 1584) it is tallied without
  // tracing, in the column where the straight line leaves the domain.  Null: unused.
  const float* leUB;
  float limMax;  // largest first-stage limit of the roulette any event can get, -log(zetaMin / (pi phatMax)) or 0; < 0: unknown
  const float* dirs;  // [nDir][DIR_STRIDE]: d.x d.y d.z 1/|d.x| 1/|d.y| 1/|d.z| 4*pi*|mu| 1/(4*pi*|mu|)
  int useRayTracing, useRussianRoulette, useRRIntensity, useHybrid, numOrdersOrig, limitContrib, useSurfaceBDRF;
  int trackByComponent;
  float rouletteW, surfaceAlbedo, zetaMin, maxContrib;
  int surf_nx, surf_ny;
  const float *surf_x, *surf_y, *surf_albedo;
  SourceDev src;
  uint32_t key0, key1;
  float *fluxUp, *fluxDown, *fluxAbs, *volAbs, *intensity, *intByComp, *excess;
  // Small domains (planeParallel: 1 column, the step cloud: 32): every warp keeps a private copy of the tallies in shared
  // memory (tsmN floats; tsmOff[TAL_*] = where each tally starts in it, -1 = not staged), see warp_tally (kernels.cuh).
  int tsmN, tsmOff[5];
  // Domains of up to a few thousand columns: the tallies exist repK times (a power of two) in global memory and block b
  // adds to copy b mod repK -- copy 0 is the tally arrays themselves, copy r > 0 starts at rep + (r-1)*repStride with the
  // tallies at repOff[TAL_*] --; k_fold_replicas sums the copies in float64 after the kernel.  Fewer increments per
  // float32 element (a column of the step cloud takes 2e5 per 8 M-photon batch: 1e-4 of rounding bias) and less
  // contention per address.  repK = 1: no copies.
  float* rep;
  int repK, repStride, repOff[5];
  // fluxAbsorbed(x,y) receives exactly the increments of volumeAbsorption(x,y,:) (MCRT:644-647), so the library does not
  // tally it: it sums the column of the (raw) volume absorption after the kernel (k_abs_from_volume).  0: tally it.
  int deriveAbs;
  unsigned long long* counters;
  unsigned long long* nextPhoton;
  uint32_t* susp;  // scratch of suspended event batches: SUSP_WORDS per warp of the grid (k_transport)
  long long firstPhoton;  // photon ids of this launch are firstPhoton + [0, src.n)
};

// Compile-time specialisation of the hot loop: REG = x, y and z are all regularly spaced (the I3RC cases), so a
// ray's path length per cell is a per-ray constant and no edge arrays are read while stepping.
// FAST = the common configuration (ray tracing, a top-of-domain source, constant Lambertian albedo, original phase
// functions, no contribution limiting): the rarely used code paths are compiled out, which keeps the kernel's
// instruction footprint small (the SM's instruction cache holds ~32 KB).
// SPLIT = the field stores only the horizontally varying layers (Problem::zlut): the gather first looks the layer up.
// JUMP = the field carries empty-space codes (above) and the rays use them (regular grids, every layer stored).
// TABSM = the kernel has staged the phase-function tables of the (single) component in shared memory (tables_of).
template <bool REG, bool FAST = false, bool SPLIT = false, bool JUMP = false, bool TABSM = false>
struct ProblemT : Problem {
  static constexpr bool kRegular = REG;
  static constexpr bool kFast = FAST;
  static constexpr bool kSplit = SPLIT;
  static constexpr bool kJump = JUMP;
  static constexpr bool kTabSm = TABSM;
  static_assert(!JUMP || (REG && !SPLIT), "empty-space codes: regular grids with every layer stored");
};
// the general-purpose instantiation (probes, CPU harness): the layer table is honoured at run time
struct ProblemDyn : Problem {
  static constexpr bool kRegular = false;
  static constexpr bool kFast = false;
  static constexpr bool kSplit = true;
  static constexpr bool kJump = false;
  static constexpr bool kTabSm = false;
};
// Kernels that cross uniform slabs in one go (regular grids with a layer table): the value gathered for a cell of a uniform
// layer may be a CODE, -(1 + tau), tau = the optical path of the ray from where it is to the far side of the slab.
template <class P>
struct SlabJump {
  static constexpr bool on = P::kSplit && P::kRegular;
};
constexpr int SLAB_MIN = 2;  // fewest further layers of a slab worth a jump
// what a (possibly coded) gathered value contributes: the extinction of the cell (to be multiplied by the path length
// in it), or -- a slab code, path length 1 -- the optical path through the rest of the slab
template <class P>
I3RC_HD float ext_value(float e) {
  if (P::kJump) return fmaxf(e, 0.0f);
  if (SlabJump<P>::on) return e < 0.0f ? -e - 1.0f : e;
  return e;
}
// the table descriptors the physics reads: the integrator's (global memory), or -- TABSM kernels -- the block's staged copy
#ifdef __CUDACC__
__shared__ TableDesc g_smTables;  // (written by k_transport<.., TABSM> before any lookup)
#endif
template <class P>
I3RC_HD const TableDesc* tables_of(const P& p) {
#ifdef __CUDA_ARCH__
  if (P::kTabSm) return &g_smTables;
#endif
  return p.tables;
}
// a table element: through the read-only path from global memory, or a plain load when the table is in shared memory
template <bool SM>
I3RC_HD float tab_load(const float* q) {
#ifdef __CUDA_ARCH__
  return SM ? *q : __ldg(q);
#else
  return *q;
#endif
}

I3RC_HD int ext_index(const Problem& p, int ix, int iy, int iz) { return ix * p.esx + iy * p.esy + iz * p.esz; }
// extinction of layer iz of the column whose index (ext_index) is idx
template <class P>
I3RC_HD float ext_gather(const P& p, int idx, int iz) {
#ifdef I3RC_CHECK
  {
    const long long at = (!P::kSplit || p.nzc == 0) ? idx : (long long)(idx - iz) + (iz >= 0 && iz < p.nz ? (p.zlut[iz].x >= 0 ? p.zlut[iz].x : 0) : -1);
    if (iz < 0 || iz >= p.nz || at < 0 || at >= p.extN) {
      atomicAdd(p.counters + CNT_BAD, 1ull);
      return 0.0f;
    }
  }
#endif
  if (!P::kSplit || p.nzc == 0) return I3RC_LDG(p.ext + idx);
#ifdef __CUDA_ARCH__
  const int2 t = __ldg(p.zlut + iz);
#else
  const int2 t = p.zlut[iz];
#endif
  if (t.x < 0) {
#ifdef __CUDA_ARCH__
    return __int_as_float(t.y);
#else
    float f;
    memcpy(&f, &t.y, sizeof f);
    return f;
#endif
  }
  return I3RC_LDG(p.ext + (idx - iz) + t.x);  // (idx counts the layer with stride 1, see Problem::zlut)
}
template <class P>
I3RC_HD float ext_at(const P& p, int ix, int iy, int iz) { return ext_gather(p, ext_index(p, ix, iy, iz), iz); }

// ---- a lower bound of the optical path from a cell to the top along a direction -------------------------------------
// Regular grid, direction (ux, uy, uz) with uz > 0.  A ray that starts anywhere in cell (ix, iy, iz) crosses every layer
// above completely; in layer iz + m it lies, horizontally, within x0 + tx * [(m-1) dz, (m+1) dz] with x0 in [0, dx) (and
// likewise in y), tx = ux / uz.  Its path length in that layer is dz / uz, so the layer adds at least dz / uz times the
// smallest extinction among the cells of that footprint.  Summed over at most nLayers layers (a truncated sum is still a
// lower bound) and stopped once it exceeds `enough`.  ext is indexed [iz][iy][ix] (x fastest), periodic in x and y.
constexpr int LE_LB_LAYERS = 32;       // layers summed on domains too large for the full depth
constexpr float LE_LB_ENOUGH = 20.0f;  // no roulette budget is that large (tauFree = -log(deviate), first-stage limit < 5)
// An UPPER bound comes the same way, with the largest extinction of each footprint and with the layer the ray starts in
// (whose footprint reaches from the start point's own cell to where the ray leaves the layer); it is only valid when
// every layer up to the top has been summed, so it is INFINITY when the sum was cut short.
// The start point may be confined to a sub-box of the cell, [fx0, fx1] x [fy0, fy1] x [fz0, fz1] in fractions of the cell.
I3RC_HD void le_path_bounds(const float* ext, int nx, int ny, int nz, float dx, float dy, float dz, float ux, float uy, float uz,
                            int ix, int iy, int iz, int nLayers, float enough, float* lower, float* upper, float fx0 = 0.0f,
                            float fx1 = 1.0f, float fy0 = 0.0f, float fy1 = 1.0f, float fz0 = 0.0f, float fz1 = 1.0f) {
  *lower = INFINITY;  // (the roulette branches only see rays that leave through the top, quirk Q4)
  *upper = INFINITY;
  if (!(uz > 0.0f)) return;
  const float tx = ux / uz, ty = uy / uz, step = dz / uz;
  float lo = 0.0f, hi = 0.0f;
  bool whole = true;  // every layer up to the top has been looked at
  for (int m = 0; iz + m < nz; m++) {
    if (m > nLayers || lo > enough) {
      whole = false;
      break;
    }
    // height above the start point while the ray is in layer iz + m: between (m - fz1) dz and (m + 1 - fz0) dz (from 0 in
    // the layer it starts in)
    const float h0 = m == 0 ? 0.0f : dz * ((float)m - fz1), h1 = dz * ((float)(m + 1) - fz0);
    const float ax = tx * h0, bx = tx * h1, ay = ty * h0, by = ty * h1;
    const int cx0 = (int)floorf(fminf(ax, bx) / dx + fx0 - 1e-3f), cx1 = (int)floorf(fmaxf(ax, bx) / dx + fx1 + 1e-3f);
    const int cy0 = (int)floorf(fminf(ay, by) / dy + fy0 - 1e-3f), cy1 = (int)floorf(fmaxf(ay, by) / dy + fy1 + 1e-3f);
    float mn = INFINITY, mx = 0.0f;
    if (cx1 - cx0 + 1 >= nx || cy1 - cy0 + 1 >= ny) {
      mn = 0.0f;  // (a footprint as wide as the domain: no statement)
      mx = INFINITY;
    } else {
      const float* layer = ext + (size_t)(iz + m) * nx * ny;
      for (int cy = cy0; cy <= cy1; cy++) {
        const int jy = ((iy + cy) % ny + ny) % ny;
        for (int cx = cx0; cx <= cx1; cx++) {
          const int jx = ((ix + cx) % nx + nx) % nx;
          const float e = I3RC_LDG(layer + (size_t)jy * nx + jx);
          mn = fminf(mn, e);
          mx = fmaxf(mx, e);
        }
      }
    }
    if (m > 0) lo = fmaf(step, mn, lo);          // (the path in the layer the ray starts in may be arbitrarily short)
    hi = fmaf(step * (m == 0 ? 1.0f - fz0 : 1.0f), mx, hi);
  }
  // (float32 sums on both sides: the bounds stay clear of what the ray itself will accumulate)
  *lower = lo * (1.0f - 5e-4f) - 1e-6f;
  if (whole) *upper = hi * (1.0f + 5e-4f) + 1e-6f;
}

// ---- tallies (MCRT:513, 530, 642-649, 574-579, 662-667) ---------------------------------------------------------
// The physics functions hand their increments to a policy object: TallyNow adds at once (per-lane scheduler, probes, CPU
// harness); the transport kernel collects them (TallyLater) and commits them with the whole warp, aggregated (warp_tally).
enum { TAL_UP = 0, TAL_DOWN = 1, TAL_ABS = 2, TAL_INT = 3, TAL_VOL = 4 };
I3RC_HD float* tally_ptr(const Problem& p, int which) {
  return which == TAL_UP ? p.fluxUp : which == TAL_DOWN ? p.fluxDown : which == TAL_ABS ? p.fluxAbs : which == TAL_INT ? p.intensity : p.volAbs;
}
struct TallyNow {
  I3RC_HD void add(const Problem& p, int which, size_t off, float v) {
#ifdef __CUDA_ARCH__
    if (p.repK > 1) {
      const int r = (int)(blockIdx.x & (unsigned)(p.repK - 1));
      if (r) {
        atomicAdd(p.rep + (size_t)(r - 1) * p.repStride + p.repOff[which] + off, v);
        return;
      }
    }
#endif
    I3RC_ATOMIC_ADD(tally_ptr(p, which) + off, v);
  }
};
struct TallyLater {  // at most two increments per lane and commit point (an absorption: column + cell)
  int n, w0, w1;
  uint32_t o0, o1;
  float v0, v1;
  I3RC_HD void add(const Problem&, int w, size_t o, float x) {
    if (n == 0) {
      w0 = w;
      o0 = (uint32_t)o;
      v0 = x;
    } else {
      w1 = w;
      o1 = (uint32_t)o;
      v1 = x;
    }
    n++;
  }
};

struct Lane {
  // current ray (regular or irregular grid): linear cell index, signed index strides per axis, and the number of
  // cells left on each axis before the ray wraps around (x, y) or leaves the domain (z), counting the current one
  int idx;
  int stx, sty, stz;
  int cntx, cnty, cntz;
  float rx, ry, rz, tau, tauLimit;  // |r|: path length left to the next x/y/z cell face; optical path so far / target
  // The geometry (idx, cnt*, r*) runs ONE CELL AHEAD of the optical path: the cell whose optical path is still to be
  // added ("pending") has path length sp.  The SIGN of rx / ry / rz records whether that face was crossed when the
  // geometry left the pending cell (negative: crossed, |r| = full path length of the new cell), which is what is
  // needed to step back when the ray turns out to end inside the pending cell; cntz == 0 says the crossing left the
  // domain.  The extinctions of the pending cell and of the cell the geometry is in live in e0 / e1, which swap
  // roles on every step (dda_step<PAR>), so that a gather is issued a whole step before its value is needed and no
  // register move ever waits for it.  e = extinction of the pending cell of a ray that has stopped.
  float sp, e0, e1, e;
  int par;  // which of e0 / e1 holds the pending cell (run-time copy for callers that step one lane at a time)
  // MINUS the path length per cell along the ray (kRegular: cell width / |direction cosine|; else 1 / |cosine|, times the
  // cell's width when used).  Negative because that is the form the cell-crossing loop consumes (ray_advance: a crossed
  // face is recorded as a negative distance): the loop then keeps no second, sign-flipped copy of the three in registers.
  float nax, nay, naz;
  int done;
  int nsteps;
  int segDone;  // how the photon's last own segment ended (DONE_*), kept until its event is processed
  // photon
  float fx, fy, fz;
  int cx, cy, cz;
  float ux, uy, uz;
  float w;
  int order;
  int mode;        // what the current ray is: MODE_PHOTON (own path segment) or a local-estimate stage
  int comp, pfi;   // photon: component (0 = surface) and phase-function entry of the last event
  float eCell;     // photon: extinction of the cell of the event point
  int d;           // photon: next local-estimate direction to generate (per-lane scheduler only)
  float ev1, ev2, ev3, le2, le3;  // deviates of the current event kept between its stages (per-lane scheduler only)
  // local-estimate ray being traced (may belong to ANOTHER photon of the warp in the warp-cooperative kernel)
  int td, tcomp;
  float tcw, tcfix, ttauFree;
  int slot;  // warp-cooperative kernel: photon slot an own segment belongs to
  Rng rng;
  int active;
  uint32_t* cnt;  // event counters [CNT_N]: the warp's shared-memory block on the device, a plain array on the host
};

// ---- small helpers ------------------------------------------------------------------------------
I3RC_HD float cell_w(const float* edges, int regular, float d, int i) {
  return regular ? d : (I3RC_LDG(edges + i + 1) - I3RC_LDG(edges + i));
}
I3RC_HD float inv_abs(float d) { return fabsf(d) >= 2.0f * F_TINY ? I3RC_FDIV(1.0f, fabsf(d)) : INFINITY; }

// largest i in [0, n-1] with edges[i] <= v (edges has n+1 entries); clamps
I3RC_HD int search_edges(const float* edges, int n, float v) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (v >= I3RC_LDG(edges + mid))
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// fractional domain coordinate q in [0,1] -> (cell, offset) along one axis
I3RC_HD void locate_frac(const float* edges, int regular, int n, float e0, float emax, float q, int* i, float* f) {
  q = q - floorf(q);  // periodic (q == 1 maps to 0, as findXYIndicies does, MCRT:1368-1369)
  if (regular) {
    float g = q * (float)n;
    int k = (int)g;
    if (k > n - 1) k = n - 1;
    *i = k;
    *f = fminf(fmaxf(g - (float)k, 0.0f), 1.0f);
  } else {
    float pos = e0 + q * (emax - e0);
    int k = search_edges(edges, n, pos);
    float a = I3RC_LDG(edges + k), b = I3RC_LDG(edges + k + 1);
    *i = k;
    *f = fminf(fmaxf((pos - a) / (b - a), 0.0f), 1.0f);
  }
}
// absolute coordinate -> (cell, offset); pos is wrapped periodically when `periodic`
I3RC_HD void locate_abs(const float* edges, int regular, int n, float e0, float emax, float d, float pos, int periodic,
                        int* i, float* f) {
  if (periodic) {
    float L = emax - e0;
    pos = pos - floorf((pos - e0) / L) * L;
  }
  if (regular) {
    float g = (pos - e0) / d;
    int k = (int)g;
    if (k > n - 1) k = n - 1;
    if (k < 0) k = 0;
    *i = k;
    *f = fminf(fmaxf(g - (float)k, 0.0f), 1.0f);
  } else {
    int k = search_edges(edges, n, pos);
    float a = I3RC_LDG(edges + k), b = I3RC_LDG(edges + k + 1);
    *i = k;
    *f = fminf(fmaxf((pos - a) / (b - a), 0.0f), 1.0f);
  }
}

I3RC_HD void make_direction(float mu, float phi, float* ux, float* uy, float* uz) {  // MCRT:2041-2059
  float st = sqrtf(fmaxf(1.0f - mu * mu, 0.0f));
  float s, c;
  I3RC_SINCOS(phi, &s, &c);
  *ux = st * c;
  *uy = st * s;
  *uz = mu;
}

// computeScatteringAngle, MCRT:1390-1417 (quirk Q1 kept: the interpolation weight is not scaled)
template <bool SM = false>
I3RC_HD float scattering_angle(const float* T, int n, float xi) {
  int k = (int)(xi * (float)n);  // angleIndex - 1
  if (k + 1 < n) {
    float leftOver = xi - I3RC_FDIV((float)k, (float)n);
    return (1.0f - leftOver) * tab_load<SM>(T + k) + leftOver * tab_load<SM>(T + k + 1);
  }
  return tab_load<SM>(T + n - 1);
}

// lookUpPhaseFuncValsFromTable, MCRT:1613-1652
template <bool SM = false>
I3RC_HD float phase_lookup(const float* T, int n, float angle) {
  float deltaTheta = I3RC_FDIV(F_PI, (float)(n - 1));
  int k = (int)I3RC_FDIV(angle, deltaTheta);  // angleIndex - 1
  if (k + 1 < n) {
    float wgt = 1.0f - I3RC_FDIV(angle - (float)k * deltaTheta, deltaTheta);
    return wgt * tab_load<SM>(T + k) + (1.0f - wgt) * tab_load<SM>(T + k + 1);
  }
  return tab_load<SM>(T + n - 1);
}

// next_direct, MCRT:2086-2113 (Marchuk rotation, rejection sampling in the unit disc; one block = two rounds)
template <class P>
I3RC_HD void next_direct(const P& p, Lane& L, float cs) {
  float D = 2.0f, AX = 0.0f, AY = 0.0f;
  while (D > 1.0f) {
    float r0, r1, r2, r3;
    L.rng.next4(p.key0, p.key1, r0, r1, r2, r3);
    AX = 1.0f - 2.0f * r0;
    AY = 1.0f - 2.0f * r1;
    D = AX * AX + AY * AY;
    if (D > 1.0f) {
      AX = 1.0f - 2.0f * r2;
      AY = 1.0f - 2.0f * r3;
      D = AX * AX + AY * AY;
    }
  }
  float B = sqrtf(I3RC_FDIV(fmaxf(1.0f - cs * cs, 0.0f), D));
  AX *= B;
  AY *= B;
  B = L.ux * AX - L.uy * AY;
  D = cs - I3RC_FDIV(B, 1.0f + fabsf(L.uz));
  L.ux = L.ux * D + AX;
  L.uy = L.uy * D - AY;
  float sb = (L.uz * B >= 0.0f) ? fabsf(B) : -fabsf(B);
  L.uz = L.uz * cs - sb;
}

// ---- the ray ------------------------------------------------------------------------------------
// per-cell path length of the ray along one axis (w = width of cell i)
template <class P>
I3RC_HD float ray_dt(const P& p, float ia, const float* edges, int regular, float d, int i) {
  if (P::kRegular) return ia;
  return cell_w(edges, regular, d, i) * ia;
}

// cell indices of the ray's current cell, recovered from the per-axis counters
template <class P>
I3RC_HD int ray_ix(const P& p, const Lane& L) { return L.stx > 0 ? p.nx - L.cntx : L.cntx - 1; }
template <class P>
I3RC_HD int ray_iy(const P& p, const Lane& L) { return L.sty > 0 ? p.ny - L.cnty : L.cnty - 1; }
template <class P>
I3RC_HD int ray_iz(const P& p, const Lane& L) { return L.stz > 0 ? p.nz - L.cntz : L.cntz - 1; }  // nz / -1: outside

// Move the geometry through the cell it is in, which becomes the pending cell.  Returns true when that leaves the
// domain (no next cell to gather).  Face crossing with selects, not branches: a warp's lanes cross different faces.
template <class P>
I3RC_HD bool ray_advance(const P& p, Lane& L) {
  const float ax = fabsf(L.rx), ay = fabsf(L.ry), az = fabsf(L.rz);
  const float s = fminf(ax, fminf(ay, az));
  const bool cx = ax <= s, cy = ay <= s, cz = az <= s;
  int nx_ = L.cntx, ny_ = L.cnty, nz_ = L.cntz, idx = L.idx;
  if (cx) nx_--, idx += L.stx;
  if (cy) ny_--, idx += L.sty;
  if (cz) nz_--, idx += L.stz;
  const bool wx = nx_ == 0, wy = ny_ == 0, out = nz_ == 0;  // periodic in x and y (MCRT:1774-1788); z: MCRT:1793-1804
  if (wx) idx -= L.stx * p.nx, nx_ = p.nx;
  if (wy) idx -= L.sty * p.ny, ny_ = p.ny;
  L.idx = idx;
  L.cntx = nx_;
  L.cnty = ny_;
  L.cntz = nz_;
  const float qx = ax - s, qy = ay - s, qz = az - s;
  L.rx = cx ? ray_dt(p, L.nax, p.xe, p.xyRegular, p.dx, ray_ix(p, L)) : qx;
  L.ry = cy ? ray_dt(p, L.nay, p.ye, p.xyRegular, p.dy, ray_iy(p, L)) : qy;
  // (outside the domain there is no cell width to look up: any negative value records the crossing)
  L.rz = cz ? ((P::kRegular || !out) ? ray_dt(p, L.naz, p.ze, p.zRegular, p.dz, ray_iz(p, L)) : -1.0f) : qz;
  L.sp = s;
  return out;
}
// The same through up to n cells on every axis at once (n >= 2), all of them known to be empty: the geometry moves
// along the ray for the path length T at which the FIRST axis has crossed its n-th face (or the ray leaves the domain),
// and every axis then counts the faces it has crossed by T.  No axis moves by more than n cells, so the cell reached and
// all cells passed on the way lie within Chebyshev distance n of the starting cell.  Regular grids only.
template <class P>
I3RC_HD bool ray_advance_far(const P& p, Lane& L, int n) {
  const float ax = fabsf(L.rx), ay = fabsf(L.ry), az = fabsf(L.rz);
  const float iax = -L.nax, iay = -L.nay, iaz = -L.naz;
  const float fn = (float)(n - 1);
  float T = fminf(fmaf(fn, iax, ax), fminf(fmaf(fn, iay, ay), fmaf(fn, iaz, az)));
  T = fminf(T, fmaf((float)(L.cntz - 1), iaz, az));  // not beyond the top / the bottom of the domain
  int mx = 0, my = 0, mz = 0;
  float qx = ax - T, qy = ay - T, qz = az - T;
  if (ax <= T) {
    mx = (int)I3RC_FDIV(T - ax, iax) + 1;
    qx = fmaxf(fmaf((float)mx, iax, ax) - T, 0.0f);
  }
  if (ay <= T) {
    my = (int)I3RC_FDIV(T - ay, iay) + 1;
    qy = fmaxf(fmaf((float)my, iay, ay) - T, 0.0f);
  }
  if (az <= T) {
    mz = (int)I3RC_FDIV(T - az, iaz) + 1;
    qz = fmaxf(fmaf((float)mz, iaz, az) - T, 0.0f);
  }
  mz = mz < L.cntz ? mz : L.cntz;
  int nx_ = L.cntx - mx, ny_ = L.cnty - my, idx = L.idx + mx * L.stx + my * L.sty + mz * L.stz;
  if (nx_ <= 0) idx -= L.stx * p.nx, nx_ += p.nx;  // periodic in x and y (n < nx, ny: at most one wrap)
  if (ny_ <= 0) idx -= L.sty * p.ny, ny_ += p.ny;
  L.idx = idx;
  L.cntx = nx_;
  L.cnty = ny_;
  L.cntz -= mz;
  L.rx = qx;
  L.ry = qy;
  L.rz = qz;
  L.sp = T;
  I3RC_COUNT(L, L.mode == MODE_PHOTON ? CNT_SKIP : CNT_SKIP_LE, mx + my + mz - 1);  // cells passed without a look (one of them is the ordinary step's)
  return L.cntz == 0;
}
// The gather for the cell the geometry has just entered.  In a uniform layer with at least SLAB_MIN further layers of
// the same slab ahead (in the ray's direction), the value is a slab code: the optical path from the ray's position to the
// far side of the slab, which does not depend on the horizontal position.
template <class P>
I3RC_HD float ext_gather_ray(const P& p, const Lane& L) {
  const int iz = ray_iz(p, L);
  if (SlabJump<P>::on) {
    if (p.nzc != 0 && p.zslab) {
#ifdef __CUDA_ARCH__
      const int2 t = __ldg(p.zlut + iz);
#else
      const int2 t = p.zlut[iz];
#endif
      if (t.x < 0) {
        float e;
        memcpy(&e, &t.y, sizeof e);
#ifdef __CUDA_ARCH__
        const float4 s = __ldg(p.zslab + iz);
#else
        const float4 s = p.zslab[iz];
#endif
        const bool up = L.stz > 0;
        const int ahead = (int)(up ? s.x : s.y);
        if (ahead >= SLAB_MIN && !isinf(L.naz)) {
          // rest of this layer + the layers ahead: vertical optical depth / |cosine| (iaz = dz / |cosine| on a regular grid)
          const float tauB = fmaf(e, fabsf(L.rz), (up ? s.z : s.w) * I3RC_FDIV(-L.naz, p.dz));
          return -(tauB + 1.0f);
        }
        return e;
      }
      return I3RC_LDG(p.ext + (L.idx - iz) + t.x);
    }
  }
  return ext_gather(p, L.idx, iz);
}
// Move the geometry from where it is in a uniform layer to the far side of the slab (the ray's optical path through it
// has been taken from the slab code): every axis counts the faces crossed in the time it takes, x and y wrap as often
// as they must.  Returns true when the far side of the slab is the top or the bottom of the domain.
template <class P>
I3RC_HD bool ray_cross_slab(const P& p, Lane& L) {
  const int iz = ray_iz(p, L);
#ifdef __CUDA_ARCH__
  const float4 s = __ldg(p.zslab + iz);
#else
  const float4 s = p.zslab[iz];
#endif
  const int ahead = (int)(L.stz > 0 ? s.x : s.y);
  const float ax = fabsf(L.rx), ay = fabsf(L.ry), az = fabsf(L.rz);
  const float iax = -L.nax, iay = -L.nay, iaz = -L.naz;
  const float T = fmaf((float)ahead, iaz, az);  // path length to the far face of the last layer of the slab
  int mx = 0, my = 0;
  float qx = ax - T, qy = ay - T;
  if (ax <= T) {
    mx = (int)I3RC_FDIV(T - ax, iax) + 1;
    qx = fmaxf(fmaf((float)mx, iax, ax) - T, 0.0f);
  }
  if (ay <= T) {
    my = (int)I3RC_FDIV(T - ay, iay) + 1;
    qy = fmaxf(fmaf((float)my, iay, ay) - T, 0.0f);
  }
  int nx_ = (L.cntx - 1 - mx) % p.nx, ny_ = (L.cnty - 1 - my) % p.ny;  // cells left on the axis, minus one, modulo the period
  if (nx_ < 0) nx_ += p.nx;
  if (ny_ < 0) ny_ += p.ny;
  L.cntx = nx_ + 1;
  L.cnty = ny_ + 1;
  L.cntz -= ahead + 1;
  L.rx = qx;
  L.ry = qy;
  L.rz = iaz;  // (a whole layer ahead, if there is one)
  L.sp = 1.0f;   // the slab code is an optical path already
  I3RC_COUNT(L, L.mode == MODE_PHOTON ? CNT_SKIP : CNT_SKIP_LE, mx + my + ahead);  // cells passed without a look
  if (L.cntz <= 0) {
    L.cntz = 0;
    return true;
  }
  L.idx = ext_index(p, ray_ix(p, L), ray_iy(p, L), ray_iz(p, L));
  return false;
}
// (Re)start the pipeline from the cell the geometry is in, whose extinction is eCell: it becomes the pending cell
// (in e0: rays always start on an even step) and the gather of the next cell is issued (into e1).
template <class P>
I3RC_HD void ray_begin(const P& p, Lane& L, float eCell) {
  L.e0 = eCell;
  L.par = 0;
  if (!ray_advance(p, L)) L.e1 = ext_gather_ray(p, L);
}

// Start a ray in cell (ix,iy,iz) at offset (fx,fy,fz) inside it along direction (dx,dy,dz); ia* = 1/|cosine|;
// eCell = extinction of that cell (the caller has gathered it, ideally long before).
template <class P>
I3RC_HD void start_ray_at(const P& p, Lane& L, int ix, int iy, int iz, float fx, float fy, float fz, float dx, float dy,
                          float dz, float iax, float iay, float iaz, float tauLimit, float eCell) {
  const bool px = dx >= 0.0f, py = dy >= 0.0f, pz = dz >= 0.0f;
  L.idx = ext_index(p, ix, iy, iz);
  L.stx = px ? p.esx : -p.esx;
  L.sty = py ? p.esy : -p.esy;
  L.stz = pz ? p.esz : -p.esz;
  L.cntx = px ? p.nx - ix : ix + 1;
  L.cnty = py ? p.ny - iy : iy + 1;
  L.cntz = pz ? p.nz - iz : iz + 1;
  if (P::kRegular) {
    iax *= p.dx;
    iay *= p.dy;
    iaz *= p.dz;
  }
  L.nax = -iax;
  L.nay = -iay;
  L.naz = -iaz;
  L.rx = isinf(iax) ? INFINITY : (px ? (1.0f - fx) : fx) * ray_dt(p, iax, p.xe, p.xyRegular, p.dx, ix);
  L.ry = isinf(iay) ? INFINITY : (py ? (1.0f - fy) : fy) * ray_dt(p, iay, p.ye, p.xyRegular, p.dy, iy);
  L.rz = isinf(iaz) ? INFINITY : (pz ? (1.0f - fz) : fz) * ray_dt(p, iaz, p.ze, p.zRegular, p.dz, iz);
  L.tau = 0.0f;
  L.tauLimit = tauLimit;
  L.nsteps = 0;
  L.done = DONE_RUN;
  ray_begin(p, L, eCell);
}

// Start a ray at the event point (cx,cy,cz ; fx,fy,fz) of the lane along direction (dx,dy,dz).
template <class P>
I3RC_HD void start_ray(const P& p, Lane& L, float dx, float dy, float dz, float iax, float iay, float iaz,
                       float tauLimit) {
  start_ray_at(p, L, L.cx, L.cy, L.cz, L.fx, L.fy, L.fz, dx, dy, dz, iax, iay, iaz, tauLimit,
               ext_at(p, L.cx, L.cy, L.cz));
}

// ONE cell crossing (the body of accumulateExtinctionAlongPath's loop, MCRT:1690-1806), software-pipelined.
// The ray carries the path length left to the next face on each axis (rx, ry, rz): all quantities stay of the
// order of one cell, so the accumulated optical path does not lose precision with the distance travelled.  The cell
// is tracked as a linear index that moves by a signed stride per crossed face; a per-axis countdown replaces the
// index comparisons of the periodic wrap (x, y) and of the exit test (z).
//
// dda_step adds the optical path of the pending cell and, unless the ray ends there, moves the geometry on by one
// cell (ray_advance).  A ray that ends -- target optical path reached inside the pending cell (MCRT:1721-1731), or the
// pending cell was the last one of the domain (MCRT:1793-1804) -- just stops with DONE_STOP; which of the two it was
// and where the ray is then are worked out by ray_after_steps() / ray_stop_inside(), outside the hot loop.
// The geometry leaves the cell it is in (which becomes the pending cell); ePending, just consumed, receives the gather of
// the cell it arrives in.  With empty-space codes: a pending cell that says "the next n cells on every axis are empty"
// lets the geometry run through all of them at once; they become ONE empty pending stretch (eOther, the value in flight
// for the first of them, is replaced by a plain zero so that ITS code is not mistaken for a statement about the last).
template <class P>
I3RC_HD void ray_move_on(const P& p, Lane& L, float& ePending, float& eOther) {
  if (P::kJump && ePending < 0.0f) {  // (a code: at least -JUMP_MIN, see k_empty_code)
    const int n = (int)(-ePending);
    eOther = 0.0f;
    if (!ray_advance_far(p, L, n)) ePending = ext_gather(p, L.idx, ray_iz(p, L));
    return;
  }
  bool out;
  if (SlabJump<P>::on && eOther < 0.0f) {
    // The cell the geometry is in carries a slab code.  If the ray's optical-path limit lies beyond the slab, the
    // geometry goes to the far side at once (the next step adds the code's optical path, with path length 1: the very
    // sum tested here); if not, the ray ends in the slab and walks there cell by cell.
    if (fmaf(1.0f, -eOther - 1.0f, L.tau) <= L.tauLimit) {
      out = ray_cross_slab(p, L);
    } else {
#ifdef __CUDA_ARCH__
      eOther = __int_as_float(__ldg(p.zlut + ray_iz(p, L)).y);
#else
      memcpy(&eOther, &p.zlut[ray_iz(p, L)].y, sizeof(float));
#endif
      out = ray_advance(p, L);
    }
  } else {
    out = ray_advance(p, L);
  }
  if (!out) ePending = ext_gather_ray(p, L);  // (the register is free now: it becomes the look-ahead; one call site)
}
template <int PAR, class P>
I3RC_HD void dda_step(const P& p, Lane& L) {
  float& ePending = PAR ? L.e1 : L.e0;
  float& eOther = PAR ? L.e0 : L.e1;
  const float t = fmaf(L.sp, ext_value<P>(ePending), L.tau);
  L.nsteps++;
  if (t > L.tauLimit || L.cntz == 0) {
    L.e = ext_value<P>(ePending);
    L.done = DONE_STOP;
    return;
  }
  L.tau = t;
  ray_move_on(p, L, ePending, eOther);
}
// Two crossings in a row, the form the kernel's trace round uses: the second step is nested in the first one's
// "goes on" branch, so a running ray pays one test per step and nothing has to be re-examined in between.
template <class P>
I3RC_HD void dda_step_pair(const P& p, Lane& L) {
  const float tA = fmaf(L.sp, ext_value<P>(L.e0), L.tau);
  if (tA > L.tauLimit || L.cntz == 0) {
    L.nsteps += 1;
    L.e = ext_value<P>(L.e0);
    L.done = DONE_STOP;
    return;
  }
  L.tau = tA;
  ray_move_on(p, L, L.e0, L.e1);
  const float tB = fmaf(L.sp, ext_value<P>(L.e1), L.tau);
  L.nsteps += 2;
  if (tB > L.tauLimit || L.cntz == 0) {
    L.e = ext_value<P>(L.e1);
    L.done = DONE_STOP;
    return;
  }
  L.tau = tB;
  ray_move_on(p, L, L.e1, L.e0);
}
// one lane at a time (CPU harness, probes): the parity is kept in the lane
template <class P>
I3RC_HD void dda_step(const P& p, Lane& L) {
  if (L.par)
    dda_step<1>(p, L);
  else
    dda_step<0>(p, L);
  L.par ^= 1;
}

// After a round of steps: how a stopped ray ended, and the hang protection (a ray of more than MAX_RAY_STEPS
// crossings is dropped as "bad").
I3RC_HD void ray_after_steps(Lane& L) {
  if (L.done == DONE_STOP) {
    const float t = fmaf(L.sp, L.e, L.tau);
    if (L.cntz == 0 && !(t > L.tauLimit)) {
      L.tau = t;
      L.done = L.stz > 0 ? DONE_TOP : DONE_BOTTOM;
    } else {
      L.done = DONE_INSIDE;
    }
  }
  if (L.done == DONE_RUN && L.nsteps > MAX_RAY_STEPS) L.done = DONE_BAD;
}

// The ray stopped with DONE_INSIDE: take the geometry back to the entry of the pending cell and move it to the
// point where the target optical path is reached (MCRT:1721-1731).  Afterwards the geometry is no longer ahead:
// L.e is the extinction of the cell it is in, and ray_restart() puts a ray that goes on back into step.
template <class P>
I3RC_HD void ray_undo_to(const P& p, Lane& L, float spd) {  // spd: path length from the entry of the pending cell
  const bool mx = L.rx < 0.0f, my = L.ry < 0.0f, mz = L.rz < 0.0f;  // faces crossed when leaving the pending cell
  L.rx = (mx ? L.sp : L.rx + L.sp) - spd;
  L.ry = (my ? L.sp : L.ry + L.sp) - spd;
  L.rz = (mz ? L.sp : L.rz + L.sp) - spd;
  if (mx) L.cntx = L.cntx == p.nx ? 1 : L.cntx + 1;
  if (my) L.cnty = L.cnty == p.ny ? 1 : L.cnty + 1;
  if (mz) L.cntz += 1;
}
I3RC_HD float ray_stop_offset(const Lane& L) { return I3RC_FDIV(L.tauLimit - L.tau, L.e); }
template <class P>
I3RC_HD void ray_stop_inside(const P& p, Lane& L) {
  ray_undo_to(p, L, ray_stop_offset(L));
  L.tau = L.tauLimit;
}
template <class P>
I3RC_HD void ray_restart(const P& p, Lane& L) {
  L.idx = ext_index(p, ray_ix(p, L), ray_iy(p, L), ray_iz(p, L));
  ray_begin(p, L, L.e);
}

// offset inside the current cell of the ray's current point, per axis (keeps f where the ray does not move)
template <class P>
I3RC_HD void ray_local(const P& p, const Lane& L, float* fx, float* fy, float* fz) {
  if (!isinf(L.nax)) {
    float rem = I3RC_FDIV(fabsf(L.rx), ray_dt(p, -L.nax, p.xe, p.xyRegular, p.dx, ray_ix(p, L)));
    rem = fminf(fmaxf(rem, 0.0f), 1.0f);
    *fx = L.stx > 0 ? 1.0f - rem : rem;
  }
  if (!isinf(L.nay)) {
    float rem = I3RC_FDIV(fabsf(L.ry), ray_dt(p, -L.nay, p.ye, p.xyRegular, p.dy, ray_iy(p, L)));
    rem = fminf(fmaxf(rem, 0.0f), 1.0f);
    *fy = L.sty > 0 ? 1.0f - rem : rem;
  }
  if (L.cntz != 0 && !isinf(L.naz)) {
    float rem = I3RC_FDIV(fabsf(L.rz), ray_dt(p, -L.naz, p.ze, p.zRegular, p.dz, ray_iz(p, L)));
    rem = fminf(fmaxf(rem, 0.0f), 1.0f);
    *fz = L.stz > 0 ? 1.0f - rem : rem;
  }
}

// ---- photon life ----------------------------------------------------------------------------------
I3RC_HD float tau_of(float xi) { return -I3RC_LOG(fmaxf(F_TINY, xi)); }  // MCRT:480

I3RC_HD float abs_x(const Problem& p, int ix, float fx) {
  float a = p.xyRegular ? p.x0 + (float)ix * p.dx : I3RC_LDG(p.xe + ix);
  return a + fx * cell_w(p.xe, p.xyRegular, p.dx, ix);
}
I3RC_HD float abs_y(const Problem& p, int iy, float fy) {
  float a = p.xyRegular ? p.y0 + (float)iy * p.dy : I3RC_LDG(p.ye + iy);
  return a + fy * cell_w(p.ye, p.xyRegular, p.dy, iy);
}
I3RC_HD float abs_z(const Problem& p, int iz, float fz) {
  float a = p.zRegular ? p.z0 + (float)iz * p.dz : I3RC_LDG(p.ze + iz);
  return a + fz * cell_w(p.ze, p.zRegular, p.dz, iz);
}

// Maximum cross-section flight (MCRT:491-497, 504-511, 521-528, 586-588) from the event point to the
// next boundary or PHYSICAL collision.  Cell indices are looked up after every move (deviation Q11).
template <class P>
I3RC_HD void max_cross_section_flight(const P& p, Lane& L, float xiFirst) {
  float x = abs_x(p, L.cx, L.fx), y = abs_y(p, L.cy, L.fy), z = abs_z(p, L.cz, L.fz);
  float Lx = p.xmax - p.x0, Ly = p.ymax - p.y0;
  float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f;
  for (int it = 0;; it++) {
    if ((it & 1) == 0) L.rng.next4(p.key0, p.key1, r0, r1, r2, r3);  // (flight, acceptance) pairs: two per block
    const float xiTau = (it & 1) ? r2 : r0, xiAcc = (it & 1) ? r3 : r1;
    float dist = tau_of(it == 0 ? xiFirst : xiTau) / p.maxExt;
    x += L.ux * dist;
    y += L.uy * dist;
    z += L.uz * dist;
    if (z >= p.zmax) {
      float back = fabsf((z - p.zmax) / L.uz);
      x -= L.ux * back;
      y -= L.uy * back;
      locate_abs(p.xe, p.xyRegular, p.nx, p.x0, p.xmax, p.dx, x, 1, &L.cx, &L.fx);
      locate_abs(p.ye, p.xyRegular, p.ny, p.y0, p.ymax, p.dy, y, 1, &L.cy, &L.fy);
      L.cz = p.nz;
      L.done = DONE_TOP;
      return;
    }
    if (z <= p.z0) {
      float back = fabsf((z - p.z0) / L.uz);
      x -= L.ux * back;
      y -= L.uy * back;
      locate_abs(p.xe, p.xyRegular, p.nx, p.x0, p.xmax, p.dx, x, 1, &L.cx, &L.fx);
      locate_abs(p.ye, p.xyRegular, p.ny, p.y0, p.ymax, p.dy, y, 1, &L.cy, &L.fy);
      L.cz = -1;
      L.done = DONE_BOTTOM;
      return;
    }
    x = x - floorf((x - p.x0) / Lx) * Lx;
    y = y - floorf((y - p.y0) / Ly) * Ly;
    locate_abs(p.xe, p.xyRegular, p.nx, p.x0, p.xmax, p.dx, x, 0, &L.cx, &L.fx);
    locate_abs(p.ye, p.xyRegular, p.ny, p.y0, p.ymax, p.dy, y, 0, &L.cy, &L.fy);
    locate_abs(p.ze, p.zRegular, p.nz, p.z0, p.zmax, p.dz, z, 0, &L.cz, &L.fz);
    float e = ext_at(p, L.cx, L.cy, L.cz);
    if (xiAcc < e / p.maxExt) {
      L.done = DONE_INSIDE;
      return;
    }
    I3RC_COUNT(L, CNT_NULL, 1);
  }
}

// Begin the next path segment of the photon from its event point (MCRT:474-497); xiTau is the deviate of its
// optical path length.
template <class P>
I3RC_HD void start_segment(const P& p, Lane& L, float xiTau) {
  L.mode = MODE_PHOTON;
  if (P::kFast || p.useRayTracing) {
    start_ray(p, L, L.ux, L.uy, L.uz, inv_abs(L.ux), inv_abs(L.uy), inv_abs(L.uz), tau_of(xiTau));
  } else {
    L.nsteps = 0;
    max_cross_section_flight(p, L, xiTau);
  }
}

// Draw a photon from the source descriptor (Code/monteCarloIllumination.f95:62-424) and place it (MCRT:453-470).
// Returns the deviate of its first optical path.
template <class P>
I3RC_HD float init_photon_state(const P& p, Lane& L, long long id) {
  const SourceDev& s = p.src;
  L.rng.init((uint64_t)(p.firstPhoton + id));
  float r0, r1, r2, r3;  // the photon's first block: position, direction, first optical path (as far as it reaches)
  L.rng.next4(p.key0, p.key1, r0, r1, r2, r3);
  float qx, qy, qz = 1.0f, mu, phi, xiTau = r3;
  const float twoPi = 2.0f * F_PI;
  switch (s.kind) {
    case 1:
      qx = r0;
      qy = r1;
      mu = s.mu;
      phi = s.phi;
      break;
    case 2:
      qx = r0;
      qy = r1;
      phi = r2 * twoPi;
      mu = s.mu;
      break;
    case 3: {
      qx = r0;
      qy = r1;
      mu = -sqrtf(r2);
      phi = r3 * twoPi;
      float t1, t2, t3;
      L.rng.next4(p.key0, p.key1, xiTau, t1, t2, t3);
      break;
    }
    case 4:
      qx = s.x;
      qy = s.y;
      mu = s.mu;
      phi = s.phi;
      break;
    case 7:  // hand-filled public arrays of the reference type
#ifdef __CUDA_ARCH__
      // (read past L1: parts of the arrays may arrive while the kernel runs, and a cache line fetched for a neighbour
      //  before its own elements were there must not be served again)
      qx = __ldcg(s.ax + id);
      qy = __ldcg(s.ay + id);
      qz = __ldcg(s.az + id);
      mu = __ldcg(s.amu + id);
      phi = __ldcg(s.aphi + id);
#else
      qx = s.ax[id];
      qy = s.ay[id];
      qz = s.az[id];
      mu = s.amu[id];
      phi = s.aphi[id];
#endif
      break;
    default:
      break;
  }
  if (!P::kFast && (s.kind == 5 || s.kind == 6)) {  // internal sources
    switch (s.kind) {
      case 5: {
        qx = s.x;
        qy = s.y;
        qz = s.z;
        mu = sqrtf(r0);
        phi = r1 * twoPi;
        while (fabsf(mu) < 2.0f * F_TINY) {
          float t1, t2, t3;
          L.rng.next4(p.key0, p.key1, mu, t1, t2, t3);
          mu = sqrtf(mu);
        }
        if (!s.pointsUp) mu = -mu;
        float d0, d1, d2;
        L.rng.next4(p.key0, p.key1, d0, d1, d2, xiTau);
        if (s.hasDx) qx += s.deltaX * (1.0f - 0.5f * d0);
        if (s.hasDy) qy += s.deltaY * (1.0f - 0.5f * d1);
        break;
      }
      case 6:
        qx = s.x;
        qy = s.y;
        qz = s.z;
        mu = s.detectorMu;
        phi = s.detectorPhi;  // stored as given, like the reference (monteCarloIllumination.f95:392)
        if (s.hasDx) qx += s.deltaX * (1.0f - 0.5f * r0);
        if (s.hasDy) qy += s.deltaY * (1.0f - 0.5f * r1);
        break;
      default:
        break;
    }
  }
  locate_frac(p.xe, p.xyRegular, p.nx, p.x0, p.xmax, qx, &L.cx, &L.fx);
  locate_frac(p.ye, p.xyRegular, p.ny, p.y0, p.ymax, qy, &L.cy, &L.fy);
  if (qz >= 1.0f - 1.2e-7f) {  // 1 - spacing(1.): the top face of the top layer
    L.cz = p.nz - 1;
    L.fz = 1.0f;
  } else {
    float zz = fmaxf(qz, 0.0f);
    if (p.zRegular) {
      float g = zz * (float)p.nz;
      int k = (int)g;
      if (k > p.nz - 1) k = p.nz - 1;
      L.cz = k;
      L.fz = fminf(fmaxf(g - (float)k, 0.0f), 1.0f);
    } else {
      locate_abs(p.ze, 0, p.nz, p.z0, p.zmax, p.dz, p.z0 + zz * (p.zmax - p.z0), 0, &L.cz, &L.fz);
    }
  }
  make_direction(mu, phi, &L.ux, &L.uy, &L.uz);
  L.w = 1.0f;
  L.order = 0;
  L.active = 1;
  I3RC_COUNT(L, CNT_PHOTONS, 1);
  return xiTau;
}
template <class P>
I3RC_HD void init_photon(const P& p, Lane& L, long long id) {
  const float xiTau = init_photon_state(p, L, id);
  start_segment(p, L, xiTau);
}

// Lambertian albedo map lookup (Code/surfaceProperties.f95:121-162)
I3RC_HD float surface_reflectance(const Problem& p, float x, float y) {
  float x0 = I3RC_LDG(p.surf_x), xm = I3RC_LDG(p.surf_x + p.surf_nx);
  float y0 = I3RC_LDG(p.surf_y), ym = I3RC_LDG(p.surf_y + p.surf_ny);
  int i, j;
  float f;
  if (p.surf_nx == 1) {
    i = 0;
  } else {
    locate_abs(p.surf_x, 0, p.surf_nx, x0, xm, 0.0f, x, 1, &i, &f);
  }
  if (p.surf_ny == 1) {
    j = 0;
  } else {
    locate_abs(p.surf_y, 0, p.surf_ny, y0, ym, 0.0f, y, 1, &j, &f);
  }
  return I3RC_LDG(p.surf_albedo + j * p.surf_nx + i);
}

// One local-estimate ray as a self-contained task (40 bytes): origin, direction index, stage, and the two possible
// contribution values, so that ANY lane of the warp can trace it (warp-cooperative kernel) .
struct LeTask {
  uint32_t xy;    // ix | iy << 16
  uint32_t zdmc;  // iz | d << 16 | mode << 21 | comp << 24
  float fx, fy, fz;
  float tauLimit;
  float cw;       // weight * normalised phase function      (contribution = cw * exp(-tau), MCRT:1530,1574)
  float cfix;     // weight * zetaMin / pi                    (Iwabuchi roulette survivors, MCRT:1556,1584)
  float tauFree;  // second-stage optical path (MCRT:1576-1578)
  float e0;       // extinction of the origin cell (gathered when the task was made, so the ray can start at once)
};

// Build the local-estimate task towards direction d from the lane's event point (MCRT:1473-1510, 1540-1569); xiTau and
// xiAcc are the two deviates Iwabuchi's roulette may need.  Returns 0 when the contribution is known to be zero
// without tracing, 1 when the task has to be traced, 2 when the contribution has been worked out on the spot (a direction
// that points straight up, or a ray that is certain to survive the roulette): in t.cw, to be tallied in the column t.xy by
// tally_intensity_at.
template <class P>
I3RC_HD int make_le_task(const P& p, Lane& L, int d, float xiTau, float xiAcc, LeTask& t) {
  const float* dv = p.dirs + d * DIR_STRIDE;
  const bool vertical = ((p.vertMask >> d) & 1u) != 0;
  // What is known about the optical path to the top before anything else is looked up: for a direction that points
  // straight up the column's suffix sum above the event cell (the rest of the path is in the cell itself), otherwise
  // the lower bound of Problem::leLB.  Both loads are issued here and used further down.
  float tauFree = 0.0f, above = 0.0f, lb = 0.0f;
  const bool bounded = p.useRRIntensity && p.leLB && !vertical && p.leLBBins == 1;
  if (vertical) above = I3RC_LDG(p.colTau + (size_t)(L.cz + 1) * ((size_t)p.nx * p.ny) + (size_t)(L.cy * p.nx + L.cx));
  if (bounded) lb = I3RC_LDG(p.leLB + (((size_t)L.cz * p.ny + L.cy) * p.nx + L.cx) * p.nDir + d);  // (direction-fastest)
  if (p.useRRIntensity) {
    tauFree = tau_of(xiTau);  // MCRT:1542
    // Under the roulette no ray has a larger budget than tauFree + limMax (the first-stage limit of the largest
    // phase-function value there is); deep inside a cloud the least the ray needs is far more: no lookup at all.
    if (p.limMax >= 0.0f && (vertical ? above : lb) > tauFree + p.limMax) return 0;
  }
  float phat;
  if (L.comp < 1) {
    phat = 1.0f / F_PI;  // quirk Q10 (MCRT:1479)
  } else {
    float proj = L.ux * I3RC_LDG(dv + 0) + L.uy * I3RC_LDG(dv + 1) + L.uz * I3RC_LDG(dv + 2);
    proj = fminf(fmaxf(proj, -1.0f), 1.0f);
    float ang = acosf(proj);
    const TableDesc& T = tables_of(p)[L.comp - 1];
    const float* tab = (!P::kFast && p.useHybrid && L.order <= p.numOrdersOrig) ? T.fwdOrig : T.fwd;
    float val = phase_lookup<P::kTabSm>(tab + (size_t)L.pfi * T.nFwd, T.nFwd, ang);
    phat = val * I3RC_LDG(dv + 7);  // 1 / (4 pi |mu|), MCRT:1509
  }
  int mode = MODE_LE_PLAIN;
  float lim = INFINITY;
  if (p.useRRIntensity) {
    if (F_PI * phat <= p.zetaMin) {
      // Iwabuchi Eq 13 (MCRT:1546-1559).  The acceptance draw does not depend on the ray, so it is taken
      // first and rejected rays are never traced (the reference traces them and then discards them).
      if (!(xiAcc * p.zetaMin <= F_PI * phat)) return 0;
      mode = MODE_LE_SMALL;
      lim = tauFree;
    } else {
      mode = MODE_LE_BIG1;  // MCRT:1566-1569
      lim = -I3RC_LOG(I3RC_FDIV(p.zetaMin, fmaxf(F_TINY, F_PI * phat)));
    }
  }
  if (bounded) {
    // the budget the ray has for reaching the top, against the least and the most it can need from this cell
    const float budget = mode == MODE_LE_SMALL ? lim : lim + tauFree;
    if (lb > budget) return 0;
    const size_t cell = ((size_t)L.cz * p.ny + L.cy) * p.nx + L.cx;
    if (p.leUB && I3RC_LDG(p.leUB + cell * p.nDir + d) <= budget && (mode == MODE_LE_SMALL || lb > lim)) {
      // Certain to reach the top, and not within the first stage's limit: the contribution is the survivors' fixed one.
      // The exit column is where the straight line from the event point meets the top of the domain.
      const float h = (float)(p.nz - L.cz) - L.fz;  // layers to the top
      const float gx = (float)L.cx + L.fx + h * I3RC_FDIV(p.dz * I3RC_LDG(dv + 0), p.dx * I3RC_LDG(dv + 2));
      const float gy = (float)L.cy + L.fy + h * I3RC_FDIV(p.dz * I3RC_LDG(dv + 1), p.dy * I3RC_LDG(dv + 2));
      int ex = (int)floorf(gx) % p.nx, ey = (int)floorf(gy) % p.ny;
      if (ex < 0) ex += p.nx;
      if (ey < 0) ey += p.ny;
      t.cw = L.w * p.zetaMin * (1.0f / F_PI);
      t.xy = (uint32_t)ex | ((uint32_t)ey << 16);
      return 2;
    }
  }
  if (vertical) {
    // Straight up: the ray never leaves its column.  Same estimator, same deviates, no tracing: the optical path to the
    // top is the rest of the event cell plus the suffix sum of the layers above; the roulette stages (MCRT:1554-1559,
    // 1570-1587) only ask whether that path fits into their optical-path limits.
    const float tauTop = (1.0f - L.fz) * cell_w(p.ze, p.zRegular, p.dz, L.cz) * ext_value<P>(L.eCell) + above;
    const float cfix = L.w * p.zetaMin * (1.0f / F_PI);
    float c;
    if (mode == MODE_LE_PLAIN)
      c = L.w * phat * I3RC_EXP(-tauTop);
    else if (mode == MODE_LE_SMALL)
      c = tauTop <= lim ? cfix : 0.0f;
    else
      c = tauTop <= lim ? L.w * phat * I3RC_EXP(-tauTop) : (tauTop - lim <= tauFree ? cfix : 0.0f);
    t.cw = c;
    t.xy = (uint32_t)L.cx | ((uint32_t)L.cy << 16);
    return 2;
  }
  t.xy = (uint32_t)L.cx | ((uint32_t)L.cy << 16);
  t.zdmc = (uint32_t)L.cz | ((uint32_t)d << 16) | ((uint32_t)mode << 21) | ((uint32_t)L.comp << 24);
  t.fx = L.fx;
  t.fy = L.fy;
  t.fz = L.fz;
  t.tauLimit = lim;
  t.cw = L.w * phat;
  t.cfix = L.w * p.zetaMin * (1.0f / F_PI);
  t.tauFree = tauFree;
  t.e0 = L.eCell;
  return 1;
}

// Make the task the lane's current ray.
template <class P>
I3RC_HD void start_le_task(const P& p, Lane& L, const LeTask& t) {
  int d = (t.zdmc >> 16) & 31;
  const float* dv = p.dirs + d * DIR_STRIDE;
  L.td = d;
  L.mode = (t.zdmc >> 21) & 7;
  L.tcomp = t.zdmc >> 24;
  L.tcw = t.cw;
  L.tcfix = t.cfix;
  L.ttauFree = t.tauFree;
  start_ray_at(p, L, (int)(t.xy & 0xffff), (int)(t.xy >> 16), (int)(t.zdmc & 0xffff), t.fx, t.fy, t.fz, I3RC_LDG(dv + 0),
               I3RC_LDG(dv + 1), I3RC_LDG(dv + 2), I3RC_LDG(dv + 3), I3RC_LDG(dv + 4), I3RC_LDG(dv + 5), t.tauLimit, t.e0);
}

// contribution c of component comp towards direction d, leaving the domain through column col (MCRT:574-579, 662-667)
template <class P, class TAL>
I3RC_HD void tally_intensity_at(const P& p, Lane& L, int d, int comp, int col, float c, TAL& tal) {
  if (!P::kFast && p.limitContrib && c > p.maxContrib) {  // MCRT:1598-1609
    I3RC_ATOMIC_ADD(p.excess + comp * p.nDir + d, c - p.maxContrib);
    c = p.maxContrib;
  }
  if (c != 0.0f) {
    size_t ncol = (size_t)p.nx * p.ny;
    tal.add(p, TAL_INT, (size_t)d * ncol + col, c);
    if (!P::kFast && p.trackByComponent) I3RC_ATOMIC_ADD(p.intByComp + ((size_t)comp * p.nDir + d) * ncol + col, c);
    I3RC_COUNT(L, CNT_CONTRIB, 1);
  }
}
template <class P, class TAL>
I3RC_HD void tally_intensity(const P& p, Lane& L, float c, TAL& tal) {
  tally_intensity_at(p, L, L.td, L.tcomp, ray_iy(p, L) * p.nx + ray_ix(p, L), c, tal);
}

// A local-estimate ray has stopped (L.done != 0).  Returns 1 when the same ray goes on with its second stage
// (Iwabuchi's chained trace, MCRT:1576-1578), 0 when it is finished and its contribution has been tallied.
template <class P, class TAL>
I3RC_HD int finish_le_ray(const P& p, Lane& L, TAL& tal) {
  int done = L.done;
  L.done = DONE_RUN;
  L.nsteps = 0;  // (the caller has counted the crossings)
  float c = 0.0f;
  if (done != DONE_BAD) {
    switch (L.mode) {
      case MODE_LE_PLAIN:  // MCRT:1529-1535
        c = L.tcw * I3RC_EXP(-L.tau);
        break;
      case MODE_LE_BIG1:  // MCRT:1570-1593 (escape is detected at the TOP only: quirk Q4)
        if (done == DONE_TOP) {
          c = L.tcw * I3RC_EXP(-L.tau);
        } else if (done == DONE_INSIDE) {
          ray_stop_inside(p, L);
          L.mode = MODE_LE_BIG2;
          L.tau = 0.0f;
          L.tauLimit = L.ttauFree;
          ray_restart(p, L);
          return 1;
        }
        break;
      default:  // MODE_LE_SMALL (MCRT:1554-1559), MODE_LE_BIG2 (MCRT:1583-1587)
        if (done == DONE_TOP) c = L.tcfix;
        break;
    }
  }
  tally_intensity(p, L, c, tal);
  return 0;
}

// The photon's own path segment has stopped: record where (event point) and how, and free the ray registers.
template <class P>
I3RC_HD void segment_finished(const P& p, Lane& L) {
  L.segDone = L.done;
  L.done = DONE_RUN;
  L.nsteps = 0;  // (the caller has counted the crossings)
  if (P::kFast || p.useRayTracing) {  // (the maximum cross-section flight has written the event point itself)
    if (L.segDone == DONE_INSIDE) ray_stop_inside(p, L);
    ray_local(p, L, &L.fx, &L.fy, &L.fz);
    L.cx = ray_ix(p, L);
    L.cy = ray_iy(p, L);
    L.cz = ray_iz(p, L);
  }
}

// The photon is finished: its slot can be refilled.  The number of random deviates it consumed is counted here.
I3RC_HD void photon_done(Lane& L) {
  L.active = 0;
  I3RC_COUNT(L, CNT_RNG, L.rng.block * 4u);
}

// Boundary and collision handling of a finished segment up to (not including) the local estimate
// (MCRT:499-561, 581-649).  xi0, xi1: deviates of the event (surface: mu, phi; collision: component pick).
// Returns 1 if the photon lives on (and then wants its local estimate when computeIntensity), 0 if it is finished.
template <class P, class TAL>
I3RC_HD int photon_event(const P& p, Lane& L, float xi0, float xi1, TAL& tal) {
  const int done = L.segDone;
  if (done == DONE_BAD) {
    I3RC_COUNT(L, CNT_BAD, 1);
    photon_done(L);
    return 0;
  }
  if (done == DONE_TOP) {  // MCRT:499-514
    tal.add(p, TAL_UP, (size_t)(L.cy * p.nx + L.cx), L.w);
    I3RC_COUNT(L, CNT_TOP, 1);
    photon_done(L);
    return 0;
  }
  if (done == DONE_BOTTOM) {  // MCRT:515-580
    L.order++;
    L.cz = 0;
    L.fz = 0.0f;
    tal.add(p, TAL_DOWN, (size_t)(L.cy * p.nx + L.cx), L.w);
    I3RC_COUNT(L, CNT_SURF, 1);
    float mu = sqrtf(xi0);
    while (!(fabsf(mu) > 2.0f * F_TINY)) {  // MCRT:542-549 (needs a deviate of exactly 0: once in 2^32)
      float t1, t2, t3;
      L.rng.next4(p.key0, p.key1, mu, t1, t2, t3);
      mu = sqrtf(mu);
    }
    float phi = 2.0f * F_PI * xi1;
    if (!P::kFast && p.useSurfaceBDRF)
      L.w *= surface_reflectance(p, abs_x(p, L.cx, L.fx), abs_y(p, L.cy, L.fy));
    else
      L.w *= p.surfaceAlbedo;
    if (L.w <= F_TINY) {
      photon_done(L);
      return 0;
    }
    make_direction(mu, phi, &L.ux, &L.uy, &L.uz);
    L.comp = 0;
    L.pfi = 0;
    L.eCell = ext_at(p, L.cx, L.cy, 0);
  } else {  // collision, MCRT:581-668
    L.order++;
    I3RC_COUNT(L, CNT_COLL, 1);
    const size_t ncell = (size_t)p.nx * p.ny * p.nz;
    const size_t cell = ((size_t)L.cz * p.ny + L.cy) * p.nx + L.cx;
    L.eCell = ext_at(p, L.cx, L.cy, L.cz);  // for the rays that start here (local estimates, next segment)
    int comp = 1;
    if (p.nc > 1) {  // findIndex(xi, (/0, cumulativeExt(:)/)), MCRT:637-638
      const float xi = xi0;
      while (comp < p.nc && xi >= I3RC_LDG(p.cumExt + (size_t)(comp - 1) * ncell + cell)) comp++;
    }
    L.comp = comp;
    float ssa = I3RC_LDG(p.ssa + (size_t)(comp - 1) * ncell + cell);
    L.pfi = I3RC_LDG(p.pfIdx + (size_t)(comp - 1) * ncell + cell) - 1;
    if (L.pfi < 0) L.pfi = 0;
    if (ssa < 1.0f) {  // MCRT:642-649
      float a = L.w * (1.0f - ssa);
      if (!p.deriveAbs) tal.add(p, TAL_ABS, (size_t)(L.cy * p.nx + L.cx), a);
      tal.add(p, TAL_VOL, cell, a);
      L.w *= ssa;
      I3RC_COUNT(L, CNT_ABS, 1);
    }
  }
  return 1;
}

// After the local estimate (or directly, when no intensity is wanted): roulette and scattering (MCRT:670-688); a
// surface event just keeps its reflected direction.  xiRoulette, xiAngle: deviates of the event's block.  The photon
// may die here (L.active = 0).
template <class P>
I3RC_HD void scatter_photon(const P& p, Lane& L, float xiRoulette, float xiAngle) {
  if (L.comp >= 1) {
    if (p.useRussianRoulette && L.w < p.rouletteW * 0.5f) {  // MCRT:673-679
      if (xiRoulette * p.rouletteW >= L.w) {
        L.w = 0.0f;
        I3RC_COUNT(L, CNT_KILL, 1);
      } else {
        L.w = p.rouletteW;
      }
    }
    if (L.w <= F_TINY) {
      photon_done(L);
      return;
    }
    const TableDesc& T = tables_of(p)[L.comp - 1];
    float theta = scattering_angle<P::kTabSm>(T.inv + (size_t)L.pfi * T.nInv, T.nInv, xiAngle);
    next_direct(p, L, I3RC_COS(theta));
  }
}
// ... and the start of the next path segment (xiTau: deviate of its optical path)
template <class P>
I3RC_HD void continue_photon(const P& p, Lane& L, float xiRoulette, float xiAngle, float xiTau) {
  scatter_photon(p, L, xiRoulette, xiAngle);
  if (L.active) start_segment(p, L, xiTau);
}

// ---- per-lane scheduler (one lane owns the photon AND traces its local-estimate rays one after the other).
// Used by the single-lane CPU harness of the test-suite and by the deterministic probes; the production kernel
// (k_transport, kernels.cuh) schedules the same functions warp-cooperatively.
template <class P>
I3RC_HD void advance_le(const P& p, Lane& L) {
  while (L.d < p.nDir) {
    LeTask t;
    int d = L.d++;
    float xiTau, xiAcc;
    if ((d & 1) == 0) {  // one block serves two directions, exactly like the warp-cooperative kernel
      L.rng.next4(p.key0, p.key1, xiTau, xiAcc, L.le2, L.le3);
    } else {
      xiTau = L.le2;
      xiAcc = L.le3;
    }
    const int what = make_le_task(p, L, d, xiTau, xiAcc, t);
    if (what == 1) {
      start_le_task(p, L, t);
      return;
    }
    if (what == 2) {
      TallyNow now;
      tally_intensity_at(p, L, d, L.comp, (int)(t.xy >> 16) * p.nx + (int)(t.xy & 0xffffu), t.cw, now);
    }
  }
  continue_photon(p, L, L.ev1, L.ev2, L.ev3);
}

template <class P>
I3RC_HD void handle_event(const P& p, Lane& L) {
  I3RC_COUNT(L, L.mode == MODE_PHOTON ? CNT_CROSS_PH : CNT_CROSS_LE, L.nsteps);
  if (L.mode == MODE_PHOTON) {
    segment_finished(p, L);
    float xi0;
    L.rng.next4(p.key0, p.key1, xi0, L.ev1, L.ev2, L.ev3);  // the event's block
    TallyNow now;
    if (!photon_event(p, L, xi0, L.ev1, now)) return;
    if (p.computeIntensity) {
      L.d = 0;
      advance_le(p, L);
    } else {
      continue_photon(p, L, L.ev1, L.ev2, L.ev3);
    }
    return;
  }
  TallyNow now;
  if (finish_le_ray(p, L, now)) return;
  advance_le(p, L);
}

}  // namespace i3rc
