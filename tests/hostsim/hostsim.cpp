// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE.
// Compiles the product's __host__ __device__ transport core (csrc/transport.cuh, csrc/philox.cuh) for the
// CPU and runs it one lane at a time, so that the kernel's logic and its Philox streams can be checked
// against the oracle in this GPU-less container before GPU time is spent.  It is not part of the product
// and is never loaded by the package: the shipped library only launches the __global__ kernels.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../i3rc_monte_carlo_model_b200/csrc/transport.cuh"

using namespace i3rc;

extern "C" {

struct HostSimArgs {
  int nx, ny, nz, nc;
  const float *xe, *ye, *ze;
  const float *ext, *cum, *ssa;
  const int* pf;
  int xyRegular, zRegular;
  // tables, per component
  const float* const* inv;
  const float* const* fwd;
  const float* const* fwdOrig;
  const int* nInv;
  const int* nFwd;
  const int* nEntries;
  // parameters
  int nDir;
  const float* mus;
  const float* phisDeg;
  int useRayTracing, useRussianRoulette, useRRIntensity, useHybrid, numOrdersOrig, limitContrib, trackByComponent;
  float surfaceAlbedo, zetaMin, maxContrib;
  int surf_nx, surf_ny;
  const float *surf_x, *surf_y, *surf_p;
  // source
  int kind;
  long long n;
  float solarMu, solarAzimuthDeg, sx, sy, sz, detectorMu, detectorPhi;
  int pointsUp, hasDx, hasDy;
  float deltaX, deltaY;
  const float *ax, *ay, *az, *amu, *aphi;
  uint32_t key0, key1;
  // outputs (raw, un-normalised tallies)
  float *fluxUp, *fluxDown, *fluxAbs, *volAbs, *intensity, *intByComp, *excess;
  unsigned long long* counters;  // [CNT_N]
};

void hostsim_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  u32x4 c = {c0, c1, c2, c3};
  u32x4 r = philox4x32_10(c, k0, k1);
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
  out[3] = r.w;
}

int hostsim_trace_rays(const HostSimArgs* a, int n, const float* pos, const float* dir, const float* tauLimit,
                       float* tauOut, float* posOut, int* idxOut);

static void fill(const HostSimArgs* a, Problem& p, std::vector<TableDesc>& td, std::vector<float>& dirs) {
  memset(&p, 0, sizeof(p));
  p.nx = a->nx; p.ny = a->ny; p.nz = a->nz; p.nc = a->nc;
  p.xyRegular = a->xyRegular; p.zRegular = a->zRegular;
  p.xe = a->xe; p.ye = a->ye; p.ze = a->ze;
  p.x0 = a->xe[0]; p.y0 = a->ye[0]; p.z0 = a->ze[0];
  p.xmax = a->xe[a->nx]; p.ymax = a->ye[a->ny]; p.zmax = a->ze[a->nz];
  p.dx = a->xe[1] - a->xe[0]; p.dy = a->ye[1] - a->ye[0]; p.dz = a->ze[1] - a->ze[0];
  p.ext = a->ext; p.esx = 1; p.esy = a->nx; p.esz = a->nx * a->ny; p.zlut = nullptr; p.nzc = 0; p.extN = (long long)a->nx * a->ny * a->nz; p.cumExt = a->cum; p.ssa = a->ssa; p.pfIdx = a->pf;
  size_t ncell = (size_t)a->nx * a->ny * a->nz;
  float mx = 0.0f;
  for (size_t i = 0; i < ncell; i++) mx = a->ext[i] > mx ? a->ext[i] : mx;
  p.maxExt = mx;
  td.resize(a->nc);
  for (int c = 0; c < a->nc; c++) {
    td[c].inv = a->inv[c]; td[c].fwd = a->fwd ? a->fwd[c] : nullptr;
    td[c].fwdOrig = a->fwdOrig ? a->fwdOrig[c] : td[c].fwd;
    td[c].nInv = a->nInv[c]; td[c].nFwd = a->nFwd ? a->nFwd[c] : 0; td[c].nEntries = a->nEntries[c]; td[c].pad = 0;
  }
  p.tables = td.data();
  p.computeIntensity = a->nDir > 0; p.nDir = a->nDir;
  dirs.assign((size_t)a->nDir * DIR_STRIDE + 1, 0.0f);
  for (int i = 0; i < a->nDir; i++) {
    float mu = a->mus[i], phi = a->phisDeg[i] * F_PI / 180.0f;
    float st = sqrtf(1.0f - mu * mu);
    float* d = &dirs[(size_t)i * DIR_STRIDE];
    d[0] = st * cosf(phi); d[1] = st * sinf(phi); d[2] = mu;
    for (int k = 0; k < 3; k++) d[3 + k] = fabsf(d[k]) >= 2.0f * F_TINY ? 1.0f / fabsf(d[k]) : INFINITY;
    d[6] = 4.0f * F_PI * fabsf(mu);
    d[7] = 1.0f / d[6];
  }
  p.dirs = dirs.data();
  p.useRayTracing = a->useRayTracing; p.useRussianRoulette = a->useRussianRoulette; p.useRRIntensity = a->useRRIntensity;
  p.useHybrid = a->useHybrid; p.numOrdersOrig = a->numOrdersOrig; p.limitContrib = a->limitContrib;
  p.trackByComponent = a->trackByComponent || a->limitContrib;
  p.useSurfaceBDRF = a->surf_nx > 0;
  p.rouletteW = 1.0f; p.surfaceAlbedo = a->surfaceAlbedo; p.zetaMin = a->zetaMin; p.maxContrib = a->maxContrib;
  p.surf_nx = a->surf_nx; p.surf_ny = a->surf_ny; p.surf_x = a->surf_x; p.surf_y = a->surf_y; p.surf_albedo = a->surf_p;
  SourceDev& s = p.src;
  s.kind = a->kind; s.n = a->n; s.mu = -fabsf(a->solarMu); s.phi = a->solarAzimuthDeg * acosf(-1.0f) / 180.0f;
  s.x = a->sx; s.y = a->sy; s.z = a->sz; s.detectorMu = a->detectorMu; s.detectorPhi = a->detectorPhi;
  s.pointsUp = a->pointsUp; s.hasDx = a->hasDx; s.hasDy = a->hasDy; s.deltaX = a->deltaX; s.deltaY = a->deltaY;
  s.ax = a->ax; s.ay = a->ay; s.az = a->az; s.amu = a->amu; s.aphi = a->aphi;
  p.key0 = a->key0; p.key1 = a->key1;
  p.fluxUp = a->fluxUp; p.fluxDown = a->fluxDown; p.fluxAbs = a->fluxAbs; p.volAbs = a->volAbs;
  p.intensity = a->intensity; p.intByComp = a->intByComp; p.excess = a->excess;
  p.counters = a->counters; p.nextPhoton = nullptr; p.firstPhoton = 0;
  p.limMax = -1.0f;  // (unknown: no early exit before the phase-function lookup)
}

}  // extern "C"

// Empty-space codes on the CPU (what k_empty_init / k_empty_grow / k_empty_code do on the device), for a field in any
// layout given by its strides: Chebyshev distance to the nearest cell with extinction, periodic in x and y.
static int g_jump = 0;
static int g_lb = 0;  // bounds of the optical path to the top (Problem::leLB / leUB) built on the CPU like k_le_path_bounds
static std::vector<float> g_leLB, g_leUB;   // g_lb: 1 = lower bound only, 2 = lower and upper bound
static void lower_bound_setup(Problem& p) {
  p.leLB = nullptr;
  p.leUB = nullptr;
  p.leLBBins = 1;
  if (!g_lb || !p.computeIntensity || !p.useRRIntensity || !(p.xyRegular && p.zRegular)) return;
  const size_t ncell = (size_t)p.nx * p.ny * p.nz;
  g_leLB.assign(ncell * p.nDir, 0.0f);
  g_leUB.assign(ncell * p.nDir, INFINITY);
  for (int d = 0; d < p.nDir; d++)
    for (size_t i = 0; i < ncell; i++) {
      const int ix = (int)(i % p.nx), iy = (int)((i / p.nx) % p.ny), iz = (int)(i / ((size_t)p.nx * p.ny));
      if (!(p.ext[i] > 0.0f || iz == 0)) continue;
      le_path_bounds(p.ext, p.nx, p.ny, p.nz, p.dx, p.dy, p.dz, p.dirs[d * DIR_STRIDE], p.dirs[d * DIR_STRIDE + 1],
                     p.dirs[d * DIR_STRIDE + 2], ix, iy, iz, p.nz, LE_LB_ENOUGH, &g_leLB[i * p.nDir + d],
                     &g_leUB[i * p.nDir + d]);
    }
  p.leLB = g_leLB.data();
  if (g_lb == 2) p.leUB = g_leUB.data();
  // the largest first-stage limit any event can get (Problem::limMax), like fill_problem of the library
  float mx = 0.0f, inv = 0.0f;
  for (int c = 0; c < p.nc; c++)
    for (const float* tab : {p.tables[c].fwd, p.tables[c].fwdOrig})
      if (tab)
        for (size_t i = 0; i < (size_t)p.tables[c].nFwd * p.tables[c].nEntries; i++) mx = mx > tab[i] ? mx : tab[i];
  for (int d = 0; d < p.nDir; d++) inv = inv > p.dirs[d * DIR_STRIDE + 7] ? inv : p.dirs[d * DIR_STRIDE + 7];
  const float phatMax = mx * inv;
  p.limMax = F_PI * phatMax > p.zetaMin ? -logf(p.zetaMin / (F_PI * phatMax)) * (1.0f + 1e-5f) + 1e-5f : 0.0f;
}
static int g_vertical = 0;  // straight-up radiance directions from column suffix sums (Problem::colTau) instead of traced
static std::vector<float> g_colTau;
static void vertical_setup(Problem& p) {
  p.colTau = nullptr;
  p.vertMask = 0;
  if (!g_vertical) return;
  const size_t ncol = (size_t)p.nx * p.ny;
  g_colTau.assign(ncol * (p.nz + 1), 0.0f);
  for (size_t c = 0; c < ncol; c++) {
    float run = 0.0f;
    for (int k = p.nz - 1; k >= 0; k--) {
      run = fmaf(p.ext[(size_t)k * ncol + c], p.ze[k + 1] - p.ze[k], run);  // (hostsim's field is x fastest, like d_ext)
      g_colTau[(size_t)k * ncol + c] = run;
    }
  }
  p.colTau = g_colTau.data();
  for (int d = 0; d < p.nDir && d < 32; d++)
    if (p.dirs[d * DIR_STRIDE] == 0.0f && p.dirs[d * DIR_STRIDE + 1] == 0.0f && p.dirs[d * DIR_STRIDE + 2] == 1.0f) p.vertMask |= 1u << d;
}
static std::vector<float> coded_field(const Problem& p) {
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  const size_t n = (size_t)nx * ny * nz;
  std::vector<float> out(p.ext, p.ext + n);
  std::vector<uint8_t> D(n);
  auto at = [&](int ix, int iy, int iz) { return (size_t)ext_index(p, ix, iy, iz); };
  for (size_t i = 0; i < n; i++) D[i] = p.ext[i] > 0.0f ? 0 : 255;
  for (int k = 1; k <= JUMP_MAX + 2; k++)
    for (int ix = 0; ix < nx; ix++)
      for (int iy = 0; iy < ny; iy++)
        for (int iz = 0; iz < nz; iz++) {
          if (D[at(ix, iy, iz)] != 255) continue;
          bool hit = false;
          for (int dx = -1; dx <= 1 && !hit; dx++)
            for (int dy = -1; dy <= 1 && !hit; dy++)
              for (int dz = -1; dz <= 1 && !hit; dz++) {
                const int jz = iz + dz;
                if (jz < 0 || jz >= nz) continue;
                hit = D[at((ix + dx + nx) % nx, (iy + dy + ny) % ny, jz)] == k - 1;
              }
          if (hit) D[at(ix, iy, iz)] = (uint8_t)k;
        }
  for (size_t i = 0; i < n; i++)
    if (D[i] >= JUMP_MIN + 2) out[i] = -(float)(D[i] - 2 < JUMP_MAX ? D[i] - 2 : JUMP_MAX);
  return out;
}

template <bool REG, bool JUMP = false>
static void run_lanes(const HostSimArgs* a, const Problem& p0) {
  ProblemT<REG, false, false, JUMP> p;
  static_cast<Problem&>(p) = p0;
  std::vector<float> coded;
  if (JUMP) {
    coded = coded_field(p0);
    p.ext = coded.data();
  }
  Lane L;
  memset(&L, 0, sizeof(L));
  uint32_t cnt[CNT_N] = {0};
  L.cnt = cnt;
  for (long long id = 0; id < a->n; id++) {
    L.active = 0; L.done = DONE_RUN; L.mode = MODE_PHOTON;
    init_photon(p, L, id);
    while (L.active) {
      if (L.done == DONE_RUN) { dda_step(p, L); ray_after_steps(L); }
      else handle_event(p, L);
    }
  }
  for (int i = 0; i < CNT_N; i++) a->counters[i] += cnt[i];
}

extern "C" {

int hostsim_run(const HostSimArgs* a) {
  Problem p;
  std::vector<TableDesc> td;
  std::vector<float> dirs;
  fill(a, p, td, dirs);
  vertical_setup(p);
  lower_bound_setup(p);
  // the same specialisation rule as the product's launcher (api.cu)
  if (p.xyRegular && p.zRegular && g_jump && p.useRayTracing) run_lanes<true, true>(a, p);
  else if (p.xyRegular && p.zRegular) run_lanes<true>(a, p);
  else run_lanes<false>(a, p);
  return 0;
}

void hostsim_set_jump(int on) { g_jump = on; }
void hostsim_set_vertical(int on) { g_vertical = on; }
void hostsim_set_lower_bound(int on) { g_lb = on; }
}  // extern "C"

template <class P>
static void trace_rays_t(P& p, int n, const float* pos, const float* dir, const float* tauLimit, float* tauOut, float* posOut,
                         int* idxOut, unsigned long long* skipped) {
  for (int r = 0; r < n; r++) {
    Lane L;
    memset(&L, 0, sizeof(L));
    uint32_t cnt[CNT_N] = {0};
    L.cnt = cnt;
    locate_abs(p.xe, p.xyRegular, p.nx, p.x0, p.xmax, p.dx, pos[3 * r], 1, &L.cx, &L.fx);
    locate_abs(p.ye, p.xyRegular, p.ny, p.y0, p.ymax, p.dy, pos[3 * r + 1], 1, &L.cy, &L.fy);
    locate_abs(p.ze, p.zRegular, p.nz, p.z0, p.zmax, p.dz, pos[3 * r + 2], 0, &L.cz, &L.fz);
    float dx = dir[3 * r], dy = dir[3 * r + 1], dz = dir[3 * r + 2];
    start_ray(p, L, dx, dy, dz, inv_abs(dx), inv_abs(dy), inv_abs(dz), tauLimit ? tauLimit[r] : INFINITY);
    while (L.done == DONE_RUN) { dda_step(p, L); ray_after_steps(L); }
    if (L.done == DONE_INSIDE) ray_stop_inside(p, L);
    tauOut[r] = L.done == DONE_BAD ? -2.0f : L.tau;
    ray_local(p, L, &L.fx, &L.fy, &L.fz);
    const int ix = ray_ix(p, L), iy = ray_iy(p, L), iz = ray_iz(p, L);
    if (posOut) {
      posOut[3 * r] = abs_x(p, ix, L.fx);
      posOut[3 * r + 1] = abs_y(p, iy, L.fy);
      posOut[3 * r + 2] = L.done == DONE_TOP ? p.zmax : (L.done == DONE_BOTTOM ? p.z0 : abs_z(p, iz, L.fz));
    }
    if (idxOut) { idxOut[3 * r] = ix + 1; idxOut[3 * r + 1] = iy + 1; idxOut[3 * r + 2] = iz + 1; }
    if (skipped) skipped[0] += cnt[CNT_SKIP], skipped[1] += (unsigned long long)L.nsteps;
  }
}

// The layer-compacted gather field of the library (Problem::zlut / zslab: horizontally uniform layers kept out of the 3-D
// array, runs of them crossed in one go) built on the CPU, to run the regular-grid SPLIT stepping code one lane at a time.
struct SplitField {
  std::vector<float> ext;
  std::vector<int2> lut;
  std::vector<float4> slab;
  int nzc = 0;
};
static SplitField split_field(const Problem& p) {
  SplitField f;
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  const size_t ncol = (size_t)nx * ny;
  f.lut.resize(nz);
  std::vector<int> layers;
  std::vector<float> uni(nz, 0.0f);
  for (int k = 0; k < nz; k++) {
    const float* row = p.ext + (size_t)k * ncol;  // (hostsim's field is x fastest: [z][y][x])
    bool same = true;
    for (size_t c = 1; c < ncol && same; c++) same = row[c] == row[0];
    if (same) {
      int bits;
      memcpy(&bits, &row[0], sizeof bits);
      f.lut[k] = int2{-1, bits};
      uni[k] = row[0];
    } else {
      f.lut[k] = int2{(int)layers.size(), 0};
      layers.push_back(k);
    }
  }
  f.nzc = (int)layers.size();
  f.ext.assign(ncol * (f.nzc ? f.nzc : 1), 0.0f);
  for (int ix = 0; ix < nx; ix++)
    for (int iy = 0; iy < ny; iy++)
      for (int k = 0; k < f.nzc; k++) f.ext[((size_t)ix * ny + iy) * f.nzc + k] = p.ext[((size_t)layers[k] * ny + iy) * nx + ix];
  f.slab.assign(nz, float4{0.f, 0.f, 0.f, 0.f});
  for (int k = 0; k < nz; k++) {
    if (f.lut[k].x >= 0) continue;
    int up = 0, dn = 0;
    float vUp = 0.0f, vDn = 0.0f;
    for (int j = k + 1; j < nz && f.lut[j].x < 0; j++) up++, vUp += uni[j] * (p.ze[j + 1] - p.ze[j]);
    for (int j = k - 1; j >= 0 && f.lut[j].x < 0; j--) dn++, vDn += uni[j] * (p.ze[j + 1] - p.ze[j]);
    f.slab[k] = float4{(float)up, (float)dn, vUp, vDn};
  }
  return f;
}

extern "C" {
// the same rays through the layer-compacted field, uniform slabs crossed in one go (jump = 0: walked cell by cell)
int hostsim_trace_rays_slab(const HostSimArgs* a, int jump, int n, const float* pos, const float* dir, const float* tauLimit,
                            float* tauOut, float* posOut, int* idxOut, unsigned long long* skipped) {
  Problem p0;
  std::vector<TableDesc> td;
  std::vector<float> dirs;
  fill(a, p0, td, dirs);
  if (!(p0.xyRegular && p0.zRegular)) return 1;
  SplitField f = split_field(p0);
  if (!f.nzc) return 2;
  ProblemT<true, false, true> p;
  static_cast<Problem&>(p) = p0;
  p.ext = f.ext.data();
  p.zlut = f.lut.data();
  p.zslab = jump ? f.slab.data() : nullptr;
  p.nzc = f.nzc;
  p.esx = p.ny * f.nzc;
  p.esy = f.nzc;
  p.esz = 1;
  p.extN = (long long)p.nx * p.ny * f.nzc;
  trace_rays_t(p, n, pos, dir, tauLimit, tauOut, posOut, idxOut, skipped);
  return 0;
}
// the same rays through the field with empty-space codes (regular grids); skipped[0] += cells passed without a look,
// skipped[1] += DDA steps taken
int hostsim_trace_rays_jump(const HostSimArgs* a, int n, const float* pos, const float* dir, const float* tauLimit,
                            float* tauOut, float* posOut, int* idxOut, unsigned long long* skipped) {
  Problem p0;
  std::vector<TableDesc> td;
  std::vector<float> dirs;
  fill(a, p0, td, dirs);
  if (!(p0.xyRegular && p0.zRegular)) return 1;
  ProblemT<true, false, false, true> p;
  static_cast<Problem&>(p) = p0;
  std::vector<float> coded = coded_field(p0);
  p.ext = coded.data();
  trace_rays_t(p, n, pos, dir, tauLimit, tauOut, posOut, idxOut, skipped);
  return 0;
}

int hostsim_trace_rays(const HostSimArgs* a, int n, const float* pos, const float* dir, const float* tauLimit,
                       float* tauOut, float* posOut, int* idxOut) {
  ProblemDyn p;
  std::vector<TableDesc> td;
  std::vector<float> dirs;
  fill(a, p, td, dirs);
  trace_rays_t(p, n, pos, dir, tauLimit, tauOut, posOut, idxOut, nullptr);
  return 0;
}
}
