// C ABI of include/i3rc_b200.h: host side of the B200 integrator (handle, validation with the reference's
// status texts, device memory, kernel launches).  No CPU fallback: without a CUDA device every compute entry
// point returns I3RC_FAILURE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/i3rc_b200.h"
#include "kernels.cuh"
#include "tables.cuh"

using namespace i3rc;

namespace {

std::string g_message;  // message of failures that happen before a handle exists

inline float h_spacing(float x) {
  if (x == 0.0f) return F_TINY;
  int e;
  (void)frexpf(fabsf(x), &e);
  float s = ldexpf(1.0f, e - 24);
  return s < F_TINY ? F_TINY : s;
}

struct HostTable {
  int kind = 0, nEntries = 0, nAngles = 0;
  std::vector<int> offsets;
  std::vector<float> coefs, angles, values;
  // device copies
  int* d_offsets = nullptr;
  float *d_coefs = nullptr, *d_angles = nullptr, *d_values = nullptr;
  std::vector<int> nodesPerEntry;
  int maxNodes = 0;
};
struct DevMatrix {
  float* d = nullptr;
  int nSteps = 0, nEntries = 0;
};

}  // namespace

struct i3rc_integrator {
  int device = 0;
  cudaStream_t stream = nullptr;
  int numSMs = 148;
  // MCRT:50-142
  bool readyToCompute = false, computeIntensity = false;
  int minForwardTableSize = 9001, minInverseTableSize = 9001;
  bool useRayTracing = true, useRussianRoulette = true;
  float RussianRouletteW = 1.0f, surfaceAlbedo = 0.0f;
  bool xyRegular = false, zRegular = false;
  float deltaX = 0, deltaY = 0, deltaZ = 0;
  int nx = 0, ny = 0, nz = 0, nc = 0;
  std::vector<float> xe, ye, ze;
  float *d_xe = nullptr, *d_ye = nullptr, *d_ze = nullptr;
  float *d_ext = nullptr, *d_cum = nullptr, *d_ssa = nullptr;
  float* d_extRaw = nullptr;  // un-normalised extinction of every component (kept from the first profile swap on)
  float maxForward = -1.0f;   // largest value of any forward phase-function table in use (< 0: not known)
  float* d_leUB = nullptr;    // upper bounds (domains small enough for the full depth), Problem::leUB
  int leUpperBound = 0;       // (tuning) 1: rays that are certain to survive the roulette are tallied without tracing.  Off:
                              // measured slower (Landsat 2.44e8 against 2.48e8 photons/s; the extra code in the event
                              // batch costs more than the 6 % of the steps it saves, profiles/r02_ab_le_bounds.txt)
  int leEarlyExit = 1;        // (tuning) 0: no test against the largest possible budget before the phase-function lookup
  float* d_leLB = nullptr;    // lower bounds of the optical path to the top per radiance direction and cell (Problem::leLB)
  bool leLBValid = false;     // (rebuilt when the field or the directions change)
  int leLowerBound = 1;       // (tuning) 0: every local-estimate ray that passes the roulette is traced
  float* d_colTau = nullptr;  // column suffix sums of extinction x layer depth (Problem::colTau)
  int l2Persist = 1;          // persisting L2 access window over a gather field of 32 MB or more (512x512x256: +2.5 %,
                              // profiles/r02_ab_l2_persisting_window_les.txt)
  const void* l2WindowFor = nullptr;
  int tablesInSmem = 0;       // (experiment) stage the phase-function tables in shared memory (one 16-warp block per SM)
  int debugZeroStrides = 0;   // (experiment) every gather reads element 0 of the field: on a homogeneous domain the results
                              // are unchanged and the run time is what the kernel would need if gathers always hit L1
  int verticalShortcut = 1;   // (tuning) 0: trace straight-up local-estimate rays like all others
  float* d_extJ = nullptr;    // the gather field with empty-space codes (transport.cuh JUMP_*), when that pays
  double codedFraction = 0.0; // share of the cells that carry a code
  int skipEmpty = 0;          // (tuning / I3RC_SKIP_EMPTY) 1: build and use empty-space codes.  Off by default: measured
                              // slower on the Landsat cloud (1.80e8 .. 2.01e8 photons/s against 2.20e8, profiles/r02_summary.md):
                              // a jump costs the whole warp more than the lanes that jump save
  float* d_extZ = nullptr;  // totalExt again, z-fastest: the copy the rays gather from (see Problem::ext)
  int2* d_zlut = nullptr;   // layer table when only the horizontally varying layers are stored (Problem::zlut)
  float4* d_zslab = nullptr;  // runs of uniform layers a ray may cross in one go (Problem::zslab; regular grids)
  int slabJump = 1;           // (tuning) 0: uniform slabs are walked cell by cell
  int nzc = 0;
  int* d_pf = nullptr;
  bool absorbing = true;    // some cell of some component has ssa < 1 (else volumeAbsorption / fluxAbsorbed stay zero)
  bool volAbsDirty = true;  // d_volAbs may hold non-zero values
  std::vector<int> maxPfIndex;  // largest phaseFunctionIndex of each component (checked against its table before tracing)
  float maxExt = 0.0f;
  bool useSurfaceBDRF = false;
  int surf_nx = 0, surf_ny = 0;
  float *d_surf_x = nullptr, *d_surf_y = nullptr, *d_surf_p = nullptr;
  std::vector<float> surf_x, surf_y, surf_p;
  std::vector<HostTable> tables;
  std::vector<DevMatrix> inv, fwd, fwdOrig;
  TableDesc* d_tableDesc = nullptr;
  bool tableDescDirty = true;
  bool tableDescDirtyForMax = true;  // the forward tables changed since maxForward was taken
  int nDir = 0;
  std::vector<float> dirs;  // [nDir][DIR_STRIDE]
  float* d_dirs = nullptr;
  bool useHybrid = false;
  float hybridWidth = 7.0f;
  int numOrdersOrig = 0;
  bool useRRIntensity = false;
  float zetaMin = 0.3f;
  bool limitContrib = false;
  float maxContrib = 3.402823466e+38f;
  bool trackByComponent = false;
  // tallies
  float *d_fluxUp = nullptr, *d_fluxDown = nullptr, *d_fluxAbs = nullptr, *d_volAbs = nullptr;
  float *d_intensity = nullptr, *d_intByComp = nullptr, *d_excess = nullptr;
  unsigned long long *d_counters = nullptr, *d_next = nullptr;
  uint32_t* d_susp = nullptr;  // Problem::susp: 7 x 32 words per warp of the grid (sized at the launch)
  size_t suspN = 0;
  double* d_fold = nullptr;  // float64 sums of the tallies of a batch that is traced in pieces (run_one_batch)
  size_t foldN = 0;
  double* d_scratch = nullptr;  // slab sums
  size_t scratchN = 0;
  float* d_fscratch = nullptr;
  size_t fscratchN = 0;
  // source arrays (I3RC_SRC_ARRAYS): uploaded in two pieces on a second stream so that all but the first tenth of the
  // copy runs under the transport kernel of the first piece
  float* d_srcArrays = nullptr;
  size_t srcArraysN = 0;
  const float* hostArrays[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t copyStream = nullptr;
  std::vector<cudaEvent_t> copyDone;  // one per piece of a batch
  unsigned long long* d_avail = nullptr;   // photons of the batch uploaded so far (SourceDev::avail)
  unsigned long long* h_avail = nullptr;   // pinned: the values the copy engine writes there, one per piece
  cudaEvent_t computeDone = nullptr;
  // batch moments
  double* d_stats = nullptr;
  size_t statsN = 0;
  bool statsVolume = false;
  int statsNDir = 0;
  // timing
  std::vector<cudaEvent_t> ev;
  size_t evUsed = 0;
  double traceMs = 0.0;
  long long traceLaunches = 0, otherLaunches = 0;
  // tuning
  int blockSize = 128, blocksPerSM = 0, residentBlocks = 0, poolShape = 0, minRunning = 16, padSmem = 0, kSteps = 16;
  int eventThreshold = 4;  // (tuning) an event batch of fewer than 32 events starts only when the ring is empty and at most this many lanes trace
  int slots80 = 1;  // (tuning `slots_80`) 0: 64 photon slots per warp on domains of many columns too
  int birthLow = 4;  // (tuning) ... or when the task ring is empty and at most this many lanes are tracing
  int birthMin = 16;  // (tuning) a warp starts new photons when this many of its slots are empty (1: at once, one by one)
  float* d_rep = nullptr;   // copies of the tallies of a domain of few columns (Problem::rep)
  size_t repN = 0;
  int replicaColumns = 4096;  // (tuning) tallies are replicated in global memory for domains of at most this many columns
  int stageMaxColumns = 8;  // tallies are staged in shared memory for domains of at most this many columns
  int stageTallies = 1536;  // floats of shared memory per warp for staged tallies (few-column domains); 0 = never
  int splitLayers = 1;  // (0 never, 1 large fields, 2 always) store only the horizontally varying layers of totalExt in 3-D when that pays (Problem::zlut)
  // nccl
  void* nccl = nullptr;
  void* ncclLib = nullptr;
  std::string message;
  i3rc_counters counters{};
};

namespace {

#define CUDA_OK(h, call)                                                        \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      (h)->message = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call; \
      return I3RC_FAILURE;                                                      \
    }                                                                           \
  } while (0)

int fail(i3rc_integrator* h, const char* m) {
  h->message = m;
  return I3RC_FAILURE;
}

template <typename T>
cudaError_t upload(T** dst, const T* src, size_t n, cudaStream_t st) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  cudaError_t e = cudaMalloc((void**)dst, sizeof(T) * (n ? n : 1));
  if (e != cudaSuccess) return e;
  if (n && src) e = cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, st);
  return e;
}
template <typename T>
void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

bool have_device() {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

void free_table(HostTable& t) {
  dfree(t.d_offsets);
  dfree(t.d_coefs);
  dfree(t.d_angles);
  dfree(t.d_values);
}
void free_matrix(DevMatrix& m) {
  dfree(m.d);
  m.nSteps = m.nEntries = 0;
}

int alloc_tallies(i3rc_integrator* h) {
  size_t ncol = (size_t)h->nx * h->ny, ncell = ncol * h->nz;
  CUDA_OK(h, cudaMalloc(&h->d_fluxUp, sizeof(float) * ncol));
  CUDA_OK(h, cudaMalloc(&h->d_fluxDown, sizeof(float) * ncol));
  CUDA_OK(h, cudaMalloc(&h->d_fluxAbs, sizeof(float) * ncol));
  CUDA_OK(h, cudaMalloc(&h->d_volAbs, sizeof(float) * ncell));
  CUDA_OK(h, cudaMalloc(&h->d_counters, sizeof(unsigned long long) * CNT_N));
  CUDA_OK(h, cudaMalloc(&h->d_next, sizeof(unsigned long long)));
  CUDA_OK(h, cudaMemsetAsync(h->d_fluxUp, 0, sizeof(float) * ncol, h->stream));
  CUDA_OK(h, cudaMemsetAsync(h->d_fluxDown, 0, sizeof(float) * ncol, h->stream));
  CUDA_OK(h, cudaMemsetAsync(h->d_fluxAbs, 0, sizeof(float) * ncol, h->stream));
  CUDA_OK(h, cudaMemsetAsync(h->d_volAbs, 0, sizeof(float) * ncell, h->stream));
  CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * CNT_N, h->stream));
  return I3RC_SUCCESS;
}

int ensure_scratch(i3rc_integrator* h, size_t nDoubles) {
  if (h->scratchN < nDoubles) {
    dfree(h->d_scratch);
    CUDA_OK(h, cudaMalloc(&h->d_scratch, sizeof(double) * nDoubles));
    h->scratchN = nDoubles;
  }
  return I3RC_SUCCESS;
}
int ensure_fscratch(i3rc_integrator* h, size_t n) {
  if (h->fscratchN < n) {
    dfree(h->d_fscratch);
    CUDA_OK(h, cudaMalloc(&h->d_fscratch, sizeof(float) * n));
    h->fscratchN = n;
  }
  return I3RC_SUCCESS;
}

// common part of the two new_Integrator forms once the dense arrays are on the device (MCRT:193-254)
// From the dense totalExt / cumulativeExt on the device: the last cumulative fraction nudged to 1 + epsilon and the
// domain maximum of totalExt (MCRT:233-234), and the copy of totalExt the rays gather from (z-fastest; only the
// horizontally varying layers when that pays, see Problem::zlut).  Also run after a component profile is replaced.
struct DevTemp {  // a temporary device buffer that is freed on every path out of the scope
  void* p = nullptr;
  ~DevTemp() {
    if (p) cudaFree(p);
  }
};
int build_gather_field(i3rc_integrator* h) {
  const int nx = h->nx, ny = h->ny, nz = h->nz;
  size_t ncell = (size_t)nx * ny * nz;
  DevTemp t_max, t_mm, t_layers;
  unsigned int* d_max = nullptr;
  CUDA_OK(h, cudaMalloc(&d_max, sizeof(unsigned int)));
  t_max.p = d_max;
  CUDA_OK(h, cudaMemsetAsync(d_max, 0, sizeof(unsigned int), h->stream));
  k_bump_and_max<<<(unsigned)((ncell + 255) / 256), 256, 0, h->stream>>>(ncell, h->d_cum + (size_t)(h->nc - 1) * ncell,
                                                                        h->d_ext, d_max);
  h->otherLaunches++;
  dfree(h->d_extZ);
  dfree(h->d_zlut);
  dfree(h->d_zslab);
  h->nzc = 0;
  {
    // which layers are horizontally uniform?  If enough are, only the others are stored in 3-D (Problem::zlut)
    if (const char* e = getenv("I3RC_SPLIT_LAYERS")) h->splitLayers = atoi(e);  // (development switch: 0 never, 2 always)
    const size_t ncol = (size_t)nx * ny;
    unsigned int* d_mm = nullptr;
    CUDA_OK(h, cudaMalloc(&d_mm, sizeof(unsigned int) * 2 * nz));
    t_mm.p = d_mm;
    k_layer_minmax<<<nz, 256, 0, h->stream>>>(h->d_ext, ncol, d_mm);
    h->otherLaunches++;
    std::vector<unsigned int> mm(2 * (size_t)nz);
    CUDA_OK(h, cudaMemcpyAsync(mm.data(), d_mm, sizeof(unsigned int) * mm.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    std::vector<int2> lut(nz);
    std::vector<int> layers;
    for (int k = 0; k < nz; k++) {
      if (mm[2 * k] == mm[2 * k + 1]) {
        lut[k] = make_int2(-1, (int)mm[2 * k]);
      } else {
        lut[k] = make_int2((int)layers.size(), 0);
        layers.push_back(k);
      }
    }
    const int nzc = (int)layers.size();
    // Worth it for a field that does not stay in L2 and of which at least a quarter of the layers is uniform.  (On small
    // fields the layer table's extra dependent load costs more than the slab crossings of Problem::zslab save: LES
    // 128x128x64 5.43e7 -> 4.35e7 photons/s, radar 1.93e8 -> 1.32e8, profiles/r02_ab_uniform_slabs.txt.)
    const bool big = ncell * sizeof(float) > ((size_t)48 << 20) || h->splitLayers == 2;
    if (ncol > 1 && nzc > 0 && nzc * 4 <= nz * 3 && h->splitLayers && big) {
      int* d_layers = nullptr;
      CUDA_OK(h, upload(&d_layers, layers.data(), layers.size(), h->stream));
      t_layers.p = d_layers;
      CUDA_OK(h, upload(&h->d_zlut, lut.data(), lut.size(), h->stream));
      if (h->xyRegular && h->zRegular) {  // runs of consecutive uniform layers (Problem::zslab)
        std::vector<float4> slab(nz, make_float4(0.f, 0.f, 0.f, 0.f));
        auto layer_tau = [&](int j) {
          float e;
          memcpy(&e, &mm[2 * j], sizeof e);
          return e * (h->ze[j + 1] - h->ze[j]);
        };
        for (int k = 0; k < nz; k++) {
          if (lut[k].x >= 0) continue;
          int up = 0, dn = 0;
          float vUp = 0.0f, vDn = 0.0f;
          for (int j = k + 1; j < nz && lut[j].x < 0; j++) up++, vUp += layer_tau(j);
          for (int j = k - 1; j >= 0 && lut[j].x < 0; j--) dn++, vDn += layer_tau(j);
          slab[k] = make_float4((float)up, (float)dn, vUp, vDn);
        }
        CUDA_OK(h, upload(&h->d_zslab, slab.data(), slab.size(), h->stream));
      }
      CUDA_OK(h, cudaMalloc(&h->d_extZ, sizeof(float) * ncol * nzc));
      dim3 b(32, 8), g((unsigned)((nx + 31) / 32), (unsigned)((nzc + 31) / 32), (unsigned)ny);
      k_compact_zfast<<<g, b, 0, h->stream>>>(nx, ny, nzc, d_layers, h->d_ext, h->d_extZ);
      h->otherLaunches++;
      CUDA_OK(h, cudaStreamSynchronize(h->stream));
      h->nzc = nzc;
    } else {
      CUDA_OK(h, cudaMalloc(&h->d_extZ, sizeof(float) * ncell));
      dim3 b(32, 8), g((unsigned)((nx + 31) / 32), (unsigned)((nz + 31) / 32), (unsigned)ny);
      k_transpose_zfast<<<g, b, 0, h->stream>>>(nx, ny, nz, h->d_ext, h->d_extZ);
      h->otherLaunches++;
    }
  }
  unsigned int bits = 0;
  CUDA_OK(h, cudaMemcpyAsync(&bits, d_max, sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  memcpy(&h->maxExt, &bits, sizeof(float));
  h->leLBValid = false;
  // column suffix sums for the radiance directions that point straight up (Problem::colTau)
  dfree(h->d_colTau);
  CUDA_OK(h, cudaMalloc(&h->d_colTau, sizeof(float) * (size_t)nx * ny * (nz + 1)));
  k_column_suffix<<<(unsigned)(((size_t)nx * ny + 127) / 128), 128, 0, h->stream>>>(nz, (size_t)nx * ny, h->d_ext, h->d_ze, h->d_colTau);
  h->otherLaunches++;
  // Empty-space codes: a second copy of the gather field in which empty cells far from any extinction say how far a ray
  // may run without looking (regular grids, every layer stored, L2-resident fields wider than two maximal jumps).
  dfree(h->d_extJ);
  h->codedFraction = 0.0;
  if (const char* e = getenv("I3RC_SKIP_EMPTY")) h->skipEmpty = atoi(e);  // (development switch)
  if (h->skipEmpty && h->xyRegular && h->zRegular && h->nzc == 0 && nx > 2 * JUMP_MAX && ny > 2 * JUMP_MAX &&
      ncell * sizeof(float) <= ((size_t)48 << 20)) {
    DevTemp t_D, t_cnt;
    uint8_t* d_D = nullptr;
    unsigned long long* d_cnt = nullptr;
    CUDA_OK(h, cudaMalloc(&d_D, ncell));
    t_D.p = d_D;
    CUDA_OK(h, cudaMalloc(&d_cnt, sizeof(unsigned long long)));
    t_cnt.p = d_cnt;
    CUDA_OK(h, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), h->stream));
    CUDA_OK(h, cudaMalloc(&h->d_extJ, sizeof(float) * ncell));
    const unsigned nb = (unsigned)((ncell + 255) / 256);
    k_empty_init<<<nb, 256, 0, h->stream>>>(h->d_extZ, ncell, d_D);
    for (int k = 1; k <= JUMP_MAX + 2; k++) k_empty_grow<<<nb, 256, 0, h->stream>>>(nx, ny, nz, k, d_D);
    int jumpMin = JUMP_MIN;
    if (const char* e = getenv("I3RC_JUMP_MIN")) jumpMin = std::max(2, std::min(JUMP_MAX, atoi(e)));  // (development switch)
    k_empty_code<<<nb, 256, 0, h->stream>>>(h->d_extZ, d_D, ncell, jumpMin, h->d_extJ, d_cnt);
    h->otherLaunches += JUMP_MAX + 4;
    unsigned long long cnt = 0;
    CUDA_OK(h, cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    h->codedFraction = (double)cnt / (double)ncell;
    if (h->codedFraction < 0.02) dfree(h->d_extJ);  // hardly any empty space: the plain kernels are the faster ones
  }
  return I3RC_SUCCESS;
}

int finish_new_integrator(i3rc_integrator* h) {
  const int nx = h->nx, ny = h->ny, nz = h->nz;
  // limits of this implementation: cell indices travel as 16-bit fields, cells are addressed with 32-bit indices
  if (nx > 32767 || ny > 32767 || nz > 32767 || (size_t)nx * ny * nz > (size_t)0x7fffffff)
    return fail(h, "new_Integrator: domain too large for this implementation (at most 32767 cells per axis, 2^31-1 cells).");
  if (h->nc > 255) return fail(h, "new_Integrator: at most 255 optical components in this implementation.");
  const std::vector<float>&x = h->xe, &y = h->ye, &z = h->ze;
  float dX = x[1] - x[0], dY = y[1] - y[0], dZ = z[1] - z[0];
  bool xyReg = true, zReg = true;
  for (int i = 0; i < nx; i++)
    if (!(fabsf((x[i + 1] - x[i]) - dX) <= 2.0f * h_spacing(x[i + 1]))) xyReg = false;
  for (int i = 0; i < ny; i++)
    if (!(fabsf((y[i + 1] - y[i]) - dY) <= 2.0f * h_spacing(y[i + 1]))) xyReg = false;
  for (int i = 0; i < nz; i++)
    if (!(fabsf((z[i + 1] - z[i]) - dZ) <= h_spacing(z[i + 1]))) zReg = false;
  h->xyRegular = xyReg;
  h->zRegular = zReg;
  if (xyReg) {
    h->deltaX = dX;
    h->deltaY = dY;
  }
  if (zReg) h->deltaZ = dZ;
  CUDA_OK(h, upload(&h->d_xe, x.data(), x.size(), h->stream));
  CUDA_OK(h, upload(&h->d_ye, y.data(), y.size(), h->stream));
  CUDA_OK(h, upload(&h->d_ze, z.data(), z.size(), h->stream));
  int rcg = build_gather_field(h);
  if (rcg != I3RC_SUCCESS) return rcg;
  h->tables.resize(h->nc);
  h->inv.resize(h->nc);
  h->fwd.resize(h->nc);
  h->fwdOrig.resize(h->nc);
  int rc = alloc_tallies(h);
  if (rc != I3RC_SUCCESS) return rc;
  h->readyToCompute = true;
  h->message.clear();
  return I3RC_SUCCESS;
}

i3rc_integrator* make_handle() {
  i3rc_integrator* h = new i3rc_integrator();
  cudaGetDevice(&h->device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess) h->numSMs = prop.multiProcessorCount;
  cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  return h;
}

// the value checks of validateOpticalComponent (Code/opticalProperties.f95:966-975) on host arrays; the upper bound of the
// phase function index is checked against the component's table (maxPf is kept for that)
const char* validate_optical_arrays(const float* ext, const float* ssa, const int32_t* pf, size_t n, int* maxPf, bool* absorbs) {
  // (a 512x512x256 domain is 1e8 values to look at: a few host threads share them)
  const int nt = n > ((size_t)1 << 22) ? 8 : 1;
  std::vector<int> mx(nt, 0), flags(nt, 0);
  auto scan = [&](int t) {
    const size_t lo = n * t / nt, hi = n * (t + 1) / nt;
    bool badE = false, badS = false, badP = false, ab = false;
    int m = 0;
    for (size_t i = lo; i < hi; i++) {
      badE |= !(ext[i] >= 0.0f);
      badS |= !(ssa[i] >= 0.0f && ssa[i] <= 1.0f);
      badP |= pf[i] < 0;
      ab |= ext[i] > 0.0f && ssa[i] < 1.0f;
      m = pf[i] > m ? pf[i] : m;
    }
    mx[t] = m;
    flags[t] = (badE ? 1 : 0) | (badS ? 2 : 0) | (badP ? 4 : 0) | (ab ? 8 : 0);
  };
  if (nt == 1) {
    scan(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(scan, t);
    for (auto& x : th) x.join();
  }
  int all = 0, m = 0;
  for (int t = 0; t < nt; t++) all |= flags[t], m = std::max(m, mx[t]);
  *maxPf = m;
  *absorbs |= (all & 8) != 0;
  if (all & 1) return "validateOpticalComponent: extinction must be >= 0.";
  if (all & 2) return "validateOpticalComponent: singleScatteringAlbedo must be between 0 and 1";
  if (all & 4) return "validateOpticalComponent: phase function index is out of bounds";
  return nullptr;
}

bool valid_edges(const float* e, int n) {
  for (int i = 0; i < n; i++)
    if (!(e[i + 1] - e[i] > 0.0f)) return false;
  return true;
}

int upload_table_desc(i3rc_integrator* h) {
  if (!h->tableDescDirty && h->d_tableDesc) return I3RC_SUCCESS;
  std::vector<TableDesc> td(h->nc);
  for (int c = 0; c < h->nc; c++) {
    td[c].inv = h->inv[c].d;
    td[c].nInv = h->inv[c].nSteps;
    td[c].fwd = h->fwd[c].d;
    td[c].fwdOrig = h->fwdOrig[c].d ? h->fwdOrig[c].d : h->fwd[c].d;
    td[c].nFwd = h->fwd[c].nSteps;
    td[c].nEntries = h->inv[c].nEntries;
    td[c].pad = 0;
  }
  CUDA_OK(h, upload(&h->d_tableDesc, td.data(), td.size(), h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  h->tableDescDirty = false;
  return I3RC_SUCCESS;
}

// tabulateInversePhaseFunctions / tabulateForwardPhaseFunctions (MCRT:1809-1923)
int tabulate(i3rc_integrator* h) {
  for (int c = 0; c < h->nc; c++) {
    HostTable& t = h->tables[c];
    PhaseTableDev pd{t.kind,    t.nEntries, t.d_offsets, t.d_coefs, t.nAngles, t.d_angles,
                     t.d_values, t.maxNodes, t.nodesPerEntry.data()};
    if (!(h->inv[c].d && h->inv[c].nSteps >= h->minInverseTableSize)) {
      if (t.kind == 0) return fail(h, "tabulateInversePhaseFunctions: failed on component (no phase function table)");
      free_matrix(h->inv[c]);
      int n = h->minInverseTableSize;
      CUDA_OK(h, cudaMalloc(&h->inv[c].d, sizeof(float) * (size_t)n * t.nEntries));
      CUDA_OK(h, build_inverse_table(pd, n, h->inv[c].d, h->stream));
      h->inv[c].nSteps = n;
      h->inv[c].nEntries = t.nEntries;
      h->otherLaunches += 4;
      h->tableDescDirty = true;
  h->tableDescDirtyForMax = true;
    }
    if (h->computeIntensity && !(h->fwd[c].d && h->fwd[c].nSteps >= h->minForwardTableSize)) {
      if (t.kind == 0) return fail(h, "tabulatePhaseFunctions: failed on component (no phase function table)");
      free_matrix(h->fwd[c]);
      free_matrix(h->fwdOrig[c]);
      int n = h->minForwardTableSize;
      size_t bytes = sizeof(float) * (size_t)n * t.nEntries;
      CUDA_OK(h, cudaMalloc(&h->fwdOrig[c].d, bytes));
      CUDA_OK(h, cudaMalloc(&h->fwd[c].d, bytes));
      CUDA_OK(h, build_forward_table(pd, n, h->fwdOrig[c].d, h->stream));
      if (h->useHybrid && h->hybridWidth > 0.0f)
        CUDA_OK(h, build_hybrid_table(h->fwdOrig[c].d, h->fwd[c].d, n, t.nEntries, h->hybridWidth, h->stream));
      else
        CUDA_OK(h, cudaMemcpyAsync(h->fwd[c].d, h->fwdOrig[c].d, bytes, cudaMemcpyDeviceToDevice, h->stream));
      h->fwd[c].nSteps = h->fwdOrig[c].nSteps = n;
      h->fwd[c].nEntries = h->fwdOrig[c].nEntries = t.nEntries;
      h->otherLaunches += 3;
      h->tableDescDirty = true;
      h->tableDescDirtyForMax = true;
    }
  }
  if (h->computeIntensity && h->tableDescDirtyForMax) {  // the largest forward-table value (for Problem::limMax)
    float mx = 0.0f;
    for (int c = 0; c < h->nc; c++) {
      for (const DevMatrix* m : {&h->fwd[c], &h->fwdOrig[c]}) {
        if (!m->d) continue;
        std::vector<float> tmp((size_t)m->nSteps * m->nEntries);
        CUDA_OK(h, cudaMemcpyAsync(tmp.data(), m->d, sizeof(float) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));
        for (float v : tmp) mx = std::max(mx, v);
      }
    }
    h->maxForward = mx;
    h->tableDescDirtyForMax = false;
  }
  return upload_table_desc(h);
}

void fill_problem(i3rc_integrator* h, Problem& p) {
  memset(&p, 0, sizeof(p));
  p.nx = h->nx;
  p.ny = h->ny;
  p.nz = h->nz;
  p.nc = h->nc;
  p.xyRegular = h->xyRegular;
  p.zRegular = h->zRegular;
  p.x0 = h->xe.front();
  p.y0 = h->ye.front();
  p.z0 = h->ze.front();
  p.xmax = h->xe.back();
  p.ymax = h->ye.back();
  p.zmax = h->ze.back();
  p.dx = h->deltaX;
  p.dy = h->deltaY;
  p.dz = h->deltaZ;
  p.xe = h->d_xe;
  p.ye = h->d_ye;
  p.ze = h->d_ze;
  p.ext = h->d_extZ;
  p.zlut = h->d_zlut;
  p.zslab = h->slabJump ? h->d_zslab : nullptr;
  p.nzc = h->nzc;
  p.esx = h->ny * (h->nzc ? h->nzc : h->nz);
  p.esy = h->nzc ? h->nzc : h->nz;
  p.esz = 1;
  p.extN = (long long)h->nx * h->ny * (h->nzc ? h->nzc : h->nz);
  if (h->debugZeroStrides && !h->nzc) p.esx = p.esy = p.esz = 0;
  p.cumExt = h->d_cum;
  p.ssa = h->d_ssa;
  p.pfIdx = h->d_pf;
  p.maxExt = h->maxExt;
  p.tables = h->d_tableDesc;
  p.computeIntensity = h->computeIntensity;
  p.nDir = h->computeIntensity ? h->nDir : 0;
  p.dirs = h->d_dirs;
  p.leLB = (h->leLowerBound && h->leLBValid && h->computeIntensity && h->useRRIntensity) ? h->d_leLB : nullptr;
  p.leLBBins = 1;
  p.leUB = (p.leLB && h->leUpperBound) ? h->d_leUB : nullptr;
  p.limMax = -1.0f;
  if (h->leLowerBound && h->leEarlyExit && h->computeIntensity && h->useRRIntensity && h->maxForward >= 0.0f && h->zetaMin > 0.0f) {
    // the largest normalised phase function any direction can see: maxForward / (4 pi |mu|) with the smallest |mu|
    float inv = 0.0f;
    for (int d = 0; d < h->nDir; d++) inv = std::max(inv, h->dirs[d * DIR_STRIDE + 7]);
    const float phatMax = h->maxForward * inv;
    p.limMax = F_PI * phatMax > h->zetaMin ? -logf(h->zetaMin / (F_PI * phatMax)) * (1.0f + 1e-5f) + 1e-5f : 0.0f;
  }
  p.colTau = h->d_colTau;
  p.vertMask = 0;
  if (h->verticalShortcut && h->d_colTau)
    for (int d = 0; d < p.nDir && d < 32; d++)
      if (h->dirs[d * DIR_STRIDE] == 0.0f && h->dirs[d * DIR_STRIDE + 1] == 0.0f && h->dirs[d * DIR_STRIDE + 2] == 1.0f) p.vertMask |= 1u << d;
  p.useRayTracing = h->useRayTracing;
  p.useRussianRoulette = h->useRussianRoulette;
  p.useRRIntensity = h->useRRIntensity;
  p.useHybrid = h->useHybrid;
  p.numOrdersOrig = h->numOrdersOrig;
  p.limitContrib = h->limitContrib;
  p.useSurfaceBDRF = h->useSurfaceBDRF;
  p.trackByComponent = h->trackByComponent || h->limitContrib;
  p.rouletteW = h->RussianRouletteW;
  p.surfaceAlbedo = h->surfaceAlbedo;
  p.zetaMin = h->zetaMin;
  p.maxContrib = h->maxContrib;
  p.surf_nx = h->surf_nx;
  p.surf_ny = h->surf_ny;
  p.surf_x = h->d_surf_x;
  p.surf_y = h->d_surf_y;
  p.surf_albedo = h->d_surf_p;
  p.fluxUp = h->d_fluxUp;
  p.fluxDown = h->d_fluxDown;
  p.fluxAbs = h->d_fluxAbs;
  p.volAbs = h->d_volAbs;
  p.intensity = h->d_intensity;
  p.intByComp = h->d_intByComp;
  p.excess = h->d_excess;
  p.counters = h->d_counters;
  p.nextPhoton = h->d_next;
  p.susp = h->d_susp;
  // Tallies staged in shared memory, one private copy per warp, when the domain has so few columns that the whole GPU
  // would otherwise hammer a handful of addresses: fluxes and radiances, and the volume absorption too if it fits.
  p.deriveAbs = 1;
  p.rep = nullptr;
  p.repK = 1;
  p.repStride = 0;
  for (int& o : p.repOff) o = 0;
  p.tsmN = 0;
  for (int& o : p.tsmOff) o = -1;
  {
    const size_t ncol = (size_t)h->nx * h->ny, ncell = ncol * h->nz;
    const size_t budget = (size_t)h->stageTallies;  // floats per warp (0 switches the staging off)
    const size_t nInt = p.computeIntensity ? ncol * (size_t)p.nDir : 0;
    // (only a handful of columns: with 32 columns -- the step cloud -- plain global atomics are already the faster way,
    //  2.96e8 against 2.22e8 photons/s staged, profiles/r02_ab_staged_tallies_step_cloud.txt)
    if (3 * ncol + nInt <= budget && ncol <= (size_t)h->stageMaxColumns) {
      p.tsmOff[TAL_UP] = 0;
      p.tsmOff[TAL_DOWN] = (int)ncol;
      p.tsmOff[TAL_ABS] = (int)(2 * ncol);
      size_t n = 3 * ncol;
      if (nInt) p.tsmOff[TAL_INT] = (int)n, n += nInt;
      if (n + ncell <= budget) p.tsmOff[TAL_VOL] = (int)n, n += ncell;
      p.tsmN = (int)n;
    }
    if (p.tsmN == 0 && ncol <= (size_t)h->replicaColumns && !h->trackByComponent && !h->limitContrib) {
      // copies of the tallies in global memory: up | down | intensity | volume absorption (the absorbed flux is derived)
      const size_t stride = 2 * ncol + nInt + ncell;
      int K = 64;
      while (K > 1 && (size_t)(K - 1) * stride * sizeof(float) > ((size_t)32 << 20)) K /= 2;
      if (K > 1) {
        p.repK = K;
        p.repStride = (int)stride;
        p.repOff[TAL_UP] = 0;
        p.repOff[TAL_DOWN] = (int)ncol;
        p.repOff[TAL_ABS] = 0;  // (unused: deriveAbs)
        p.repOff[TAL_INT] = (int)(2 * ncol);
        p.repOff[TAL_VOL] = (int)(2 * ncol + nInt);
      }
    }
  }
}

// validation of a photon source: the checks of the new_PhotonStream constructors
// (Code/monteCarloIllumination.f95:78-83, 122-125, 163-164, 200-208, 251-268, 353-376)
int validate_source(i3rc_integrator* h, const i3rc_photon_source* s) {
  if (!s) return fail(h, "getNextPhoton: photons have not been initialized.");
  if (s->numberOfPhotons <= 0) return fail(h, "setIllumination: must ask for non-negative number of photons.");
  if (s->numberOfPhotons > 0x7fffffffLL)  // numberOfPhotons is a default integer in the reference; photon ids travel as 32 bits
    return fail(h, "setIllumination: at most 2^31-1 photons per batch (use more batches).");
  switch (s->kind) {
    case I3RC_SRC_DIRECTIONAL:
    case I3RC_SRC_SPOTLIGHT:
      if (s->solarAzimuth < 0.0f || s->solarAzimuth > 360.0f) return fail(h, "setIllumination: solarAzimuth out of bounds");
      /* fall through */
    case I3RC_SRC_RANDOM_AZIMUTH:
      if (fabsf(s->solarMu) > 1.0f || fabsf(s->solarMu) <= F_TINY) return fail(h, "setIllumination: solarMu out of bounds");
      break;
    case I3RC_SRC_FLUX:
      break;
    case I3RC_SRC_INTERNAL_FLUX:
    case I3RC_SRC_INTERNAL_INTENSITY:
      break;
    case I3RC_SRC_ARRAYS:
      if (!s->xPosition || !s->yPosition || !s->zPosition || !s->initialMu || !s->initialPhi)
        return fail(h, "getNextPhoton: photons have not been initialized.");
      break;
    default:
      return fail(h, "new_PhotonStream: unknown source kind");
  }
  if (s->kind == I3RC_SRC_SPOTLIGHT && (s->x > 1.0f || s->x <= 0.0f || s->y > 1.0f || s->y <= 0.0f))
    return fail(h, "setIllumination: x and y positions must be between 0 and 1");
  if (s->kind == I3RC_SRC_INTERNAL_FLUX || s->kind == I3RC_SRC_INTERNAL_INTENSITY) {
    if (s->x > 1.0f || s->x <= 0.0f || s->y > 1.0f || s->y <= 0.0f || s->z > 1.0f || s->z <= 0.0f)
      return fail(h, "setIllumination: x, y, z positions must be between 0 and 1");
    if ((s->has_deltaX && (s->x + s->deltaX / 2.0f > 1.0f || s->x - s->deltaX / 2.0f <= 0.0f)) ||
        (s->has_deltaY && (s->y + s->deltaY / 2.0f > 1.0f || s->y - s->deltaY / 2.0f <= 0.0f)))
      return fail(h, "setIllumination: max, min positions must be between 0 and 1");
  }
  if (s->kind == I3RC_SRC_INTERNAL_INTENSITY) {
    if (s->detectorPhi < 0.0f || s->detectorPhi > 360.0f) return fail(h, "setIllumination: detectorPhi out of bounds");
    if (fabsf(s->detectorMu) > 1.0f || fabsf(s->detectorMu) <= F_TINY)
      return fail(h, "setIllumination: detectorMu out of bounds");
  }
  return I3RC_SUCCESS;
}

int fill_source(i3rc_integrator* h, const i3rc_photon_source* s, SourceDev& d) {
  memset(&d, 0, sizeof(d));
  const float pi = acosf(-1.0f);
  d.kind = s->kind;
  d.n = s->numberOfPhotons;
  d.mu = -fabsf(s->solarMu);
  d.phi = s->solarAzimuth * pi / 180.0f;
  d.x = s->x;
  d.y = s->y;
  d.z = s->z;
  if (s->kind == I3RC_SRC_INTERNAL_FLUX)
    d.z = s->detectorPointsUp ? fmaxf(s->z, 2.0f * F_TINY) : fminf(s->z, 1.0f - h_spacing(1.0f));
  if (s->kind == I3RC_SRC_INTERNAL_INTENSITY)
    d.z = s->detectorMu > F_TINY ? fmaxf(s->z, 2.0f * F_TINY) : fminf(s->z, 1.0f - h_spacing(1.0f));
  d.detectorMu = s->detectorMu;
  d.detectorPhi = s->detectorPhi;
  d.pointsUp = s->detectorPointsUp;
  d.hasDx = s->has_deltaX;
  d.hasDy = s->has_deltaY;
  d.deltaX = s->deltaX;
  d.deltaY = s->deltaY;
  if (s->kind == I3RC_SRC_ARRAYS) {
    size_t n = (size_t)s->numberOfPhotons;
    if (h->srcArraysN < 5 * n) {
      dfree(h->d_srcArrays);
      CUDA_OK(h, cudaMalloc(&h->d_srcArrays, sizeof(float) * 5 * n));
      h->srcArraysN = 5 * n;
    }
    const float* srcs[5] = {s->xPosition, s->yPosition, s->zPosition, s->initialMu, s->initialPhi};
    for (int k = 0; k < 5; k++) h->hostArrays[k] = srcs[k];  // (copied by run_one_batch, piecewise)
    if (!h->copyStream) {
      CUDA_OK(h, cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
      CUDA_OK(h, cudaEventCreateWithFlags(&h->computeDone, cudaEventDisableTiming));
    }
    d.ax = h->d_srcArrays;
    d.ay = h->d_srcArrays + n;
    d.az = h->d_srcArrays + 2 * n;
    d.amu = h->d_srcArrays + 3 * n;
    d.aphi = h->d_srcArrays + 4 * n;
  }
  return I3RC_SUCCESS;
}

template <int BLOCK, bool REG, bool FAST, bool SPLIT, int MINB, int STEPS, int NSLOT, int QCAP, bool TSM = false, bool JUMP = false,
          bool TABSM = false>
int launch_transport_t(i3rc_integrator* h, const Problem& p) {
  auto kern = k_transport<BLOCK, REG, FAST, SPLIT, MINB, STEPS, NSLOT, QCAP, TSM, JUMP, TABSM>;
  size_t dynSmem = (size_t)h->padSmem + (TSM ? sizeof(float) * (BLOCK / 32) * (size_t)p.tsmN : 0);
  if (TABSM)  // the warps' state + the inverse and the forward table of the one component
    dynSmem = (sizeof(WarpShared<NSLOT, QCAP>) * (BLOCK / 32) + 15) / 16 * 16 + sizeof(float) * ((size_t)h->inv[0].nSteps + h->fwd[0].nSteps);
  if (dynSmem > 8192) CUDA_OK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem));
  int perSM = h->blocksPerSM;
  if (perSM <= 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, BLOCK, dynSmem) != cudaSuccess || perSM <= 0)
      perSM = 1;
  }
  long long want = (p.src.n + BLOCK - 1) / BLOCK;
  long long grid = (long long)h->numSMs * perSM;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  const size_t suspNeed = (size_t)grid * (BLOCK / 32) * SUSP_WORDS;
  if (h->suspN < suspNeed) {  // (scratch of suspended event batches, Problem::susp)
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    dfree(h->d_susp);
    h->suspN = std::max(suspNeed, (size_t)h->numSMs * 32 * SUSP_WORDS);
    CUDA_OK(h, cudaMalloc(&h->d_susp, sizeof(uint32_t) * h->suspN));
  }
  ProblemT<REG, FAST, SPLIT, JUMP, TABSM> pt;
  static_cast<Problem&>(pt) = p;
  pt.susp = h->d_susp;
  if (JUMP) pt.ext = h->d_extJ;  // the copy of the gather field that carries the empty-space codes
  // (the default group of births grows with the pool: 16 of 64 slots, 24 of 80 -- measured, profiles/r02_ab_builds_late.txt)
  const int birthMin = (NSLOT > 64 && h->birthMin == 16) ? 24 : h->birthMin;
  kern<<<(unsigned)grid, BLOCK, dynSmem, h->stream>>>(pt, h->eventThreshold, h->minRunning, birthMin | (h->birthLow << 8));
  return I3RC_SUCCESS;
}
// MINB = resident blocks per SM the kernel is compiled for (register cap 65536 / (MINB * BLOCK));
// STEPS = DDA crossings per bookkeeping round.  The tuning grid exists for the common configuration only (regular
// grid + the FAST feature set, see ProblemT); everything else runs the general kernel at its default shape.
template <int MINB, int NSLOT, int QCAP>
int launch_transport_fast(i3rc_integrator* h, const Problem& p) {
  const int steps = h->kSteps <= 8 ? 8 : (h->kSteps <= 16 ? 16 : 32);
  return steps == 8    ? launch_transport_t<128, true, true, false, MINB, 8, NSLOT, QCAP>(h, p)
         : steps == 16 ? launch_transport_t<128, true, true, false, MINB, 16, NSLOT, QCAP>(h, p)
                       : launch_transport_t<128, true, true, false, MINB, 32, NSLOT, QCAP>(h, p);
}
// A gather field that does not share L2 comfortably with the rest of the traffic (the per-component fields of a collision
// and the volume tallies of a 512x512x256 domain are gigabytes) is given a persisting L2 access window: its lines are
// kept, everything else streams.  (tuning `l2_persist`; set up once per gather field)
void set_l2_window(i3rc_integrator* h) {
  if (h->l2WindowFor == h->d_extZ) return;
  h->l2WindowFor = h->d_extZ;
  const size_t bytes = sizeof(float) * (size_t)h->nx * h->ny * (h->nzc ? h->nzc : h->nz);
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  int maxPersist = 0, maxWindow = 0;
  cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
  cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
  if (!h->l2Persist || bytes < ((size_t)32 << 20) || maxPersist <= 0 || maxWindow <= 0) {
    attr.accessPolicyWindow.num_bytes = 0;  // no window
    cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    return;
  }
  const size_t persist = std::min<size_t>((size_t)maxPersist, bytes);
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
  attr.accessPolicyWindow.base_ptr = h->d_extZ;
  attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)maxWindow);
  attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)attr.accessPolicyWindow.num_bytes);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
  cudaGetLastError();
}

int launch_transport(i3rc_integrator* h, const Problem& p) {
  constexpr int NS = 64;  // photon slots per warp
  set_l2_window(h);
  const bool reg = p.xyRegular && p.zRegular;
  const bool fast = p.useRayTracing && p.src.kind != 5 && p.src.kind != 6 && !p.useSurfaceBDRF && !p.useHybrid && !p.limitContrib &&
                    !p.trackByComponent;
  if (p.tsmN > 0) {  // a domain of a few columns: tallies staged per warp in shared memory (warp_tally, kernels.cuh)
    if (p.nzc) {  // (a layer table, forced by `split_layers` = 2 on a small domain: the variants that honour it)
      if (reg && fast) return launch_transport_t<128, true, true, true, 5, 16, NS, 64, true>(h, p);
      if (reg) return launch_transport_t<128, true, false, true, 5, 16, NS, 64, true>(h, p);
      return launch_transport_t<128, false, false, true, 5, 16, NS, 64, true>(h, p);
    }
    if (reg && fast)
      return h->residentBlocks == 4 ? launch_transport_t<128, true, true, false, 4, 16, NS, 64, true>(h, p)
                                    : launch_transport_t<128, true, true, false, 5, 16, NS, 64, true>(h, p);
    if (reg) return launch_transport_t<128, true, false, false, 5, 16, NS, 64, true>(h, p);
    return launch_transport_t<128, false, false, false, 5, 16, NS, 64, true>(h, p);
  }
  if (p.nzc) {  // only the horizontally varying layers are stored (large fields): the gathers look the layer up first
    if (reg && fast) {  // (5 resident blocks: 512x512x256 with slab crossings 3.86e7 photons/s against 3.63e7 with 6)
      if (h->poolShape == 4) return launch_transport_t<128, true, true, true, 5, 16, NS, 128>(h, p);  // (experiment: longer task ring)
      return h->residentBlocks == 6 ? launch_transport_t<128, true, true, true, 6, 16, NS, 64>(h, p)
                                    : launch_transport_t<128, true, true, true, 5, 16, NS, 64>(h, p);
    }
    if (reg) return launch_transport_t<128, true, false, true, 5, 16, NS, 64>(h, p);
    return launch_transport_t<128, false, false, true, 5, 16, NS, 64>(h, p);
  }
  if (reg && h->d_extJ && p.useRayTracing) {  // enough empty space for the rays to jump through it
    const int blocks = h->residentBlocks ? h->residentBlocks : 6;
    if (fast) return blocks == 5 ? launch_transport_t<128, true, true, false, 5, 16, NS, 64, false, true>(h, p)
                                 : launch_transport_t<128, true, true, false, 6, 16, NS, 64, false, true>(h, p);
    return launch_transport_t<128, true, false, false, 5, 16, NS, 64, false, true>(h, p);
  }
  // (experiment, `tables_in_smem`) one block of 16 warps per SM with the inverse and the forward phase-function table of the
  // single component staged in shared memory: measured against the default in profiles/r02_summary.md
  if (reg && fast && h->tablesInSmem && p.nc == 1 && p.computeIntensity && h->inv[0].nEntries == 1 && h->fwd[0].d &&
      sizeof(float) * ((size_t)h->inv[0].nSteps + h->fwd[0].nSteps) <= 90 * 1024)
    return launch_transport_t<512, true, true, false, 1, 16, NS, 64, false, false, true>(h, p);
  if (reg && fast) {
    // resident blocks per SM, 0 = automatic: 6 while the extinction field is L2-resident; 5 (more of the 256 KB left as
    // L1) when the gathers go to HBM
    const size_t ncell = (size_t)p.nx * p.ny * p.nz;
    const int blocks = h->residentBlocks ? h->residentBlocks : (ncell * sizeof(float) <= (size_t)48 << 20 ? 6 : 5);
    switch (blocks * 10 + h->poolShape) {  // (blocks per SM, photon slots and ring entries per warp)
      case 60:
        // Domains of many columns (long rays; the ring is not what limits the warp): 80 slots per warp, the scratch of
        // suspended batches in global memory instead -- same shared memory, same L1.  Landsat 2.71e8 -> 2.77e8 photons/s.
        // Domains of few columns keep 64: with 80 photons the ring of 64 tasks is full most of the time and the event
        // batches get suspended (step cloud 3.0e8 -> 2.4e8, profiles/r02_ab_builds_late.txt).
        // Many radiance directions fill the ring as well (16 directions on 128x128x64: 7.3e7 -> 7.1e7): 64 slots there too.
        if ((size_t)p.nx * p.ny >= 4096 && p.nDir <= 4 && h->slots80) return launch_transport_t<128, true, true, false, 6, 16, 80, 64>(h, p);
        return launch_transport_fast<6, NS, 64>(h, p);
      case 40:
        return launch_transport_fast<4, NS, 64>(h, p);
      case 52:
        return launch_transport_fast<5, 48, 64>(h, p);
      case 55:  // (experiments: more photons per warp, so that the task ring runs dry less often)
        return launch_transport_t<128, true, true, false, 5, 16, 96, 64>(h, p);
      case 56:
        return launch_transport_t<128, true, true, false, 5, 16, 80, 64>(h, p);
      case 65:
        return launch_transport_t<128, true, true, false, 6, 16, 80, 64>(h, p);
      case 72:  // (experiments: more resident warps at the price of registers)
        return launch_transport_t<128, true, true, false, 7, 16, 48, 64>(h, p);
      case 83:
        return launch_transport_t<128, true, true, false, 8, 16, 32, 64>(h, p);
      default:  // the shared-memory footprint is kept small on purpose: what is left of the 256 KB is L1 for the gathers
        return launch_transport_fast<5, NS, 64>(h, p);
    }
  }
  if (reg) return launch_transport_t<128, true, false, false, 5, 16, NS, 64>(h, p);
  return fast ? launch_transport_t<128, false, true, false, 5, 16, NS, 64>(h, p) : launch_transport_t<128, false, false, false, 5, 16, NS, 64>(h, p);
}

// zero tallies, trace one batch, post-process (MCRT:296-395); no host synchronisation
int run_one_batch(i3rc_integrator* h, const SourceDev& src, uint32_t key0, uint32_t key1) {
  size_t ncol = (size_t)h->nx * h->ny, ncell = ncol * h->nz;
  int nD = h->computeIntensity ? h->nDir : 0;
  bool byComp = h->trackByComponent || h->limitContrib;
  CUDA_OK(h, cudaMemsetAsync(h->d_fluxUp, 0, sizeof(float) * ncol, h->stream));
  CUDA_OK(h, cudaMemsetAsync(h->d_fluxDown, 0, sizeof(float) * ncol, h->stream));
  CUDA_OK(h, cudaMemsetAsync(h->d_fluxAbs, 0, sizeof(float) * ncol, h->stream));
  // (nothing is absorbed in the volume when every component has ssa = 1 everywhere: the array stays zero from the last
  // batch that could write it, and so do its normalisation and its share of the moments)
  if (h->volAbsDirty) CUDA_OK(h, cudaMemsetAsync(h->d_volAbs, 0, sizeof(float) * ncell, h->stream));
  h->volAbsDirty = h->absorbing;
  if (h->d_intensity) CUDA_OK(h, cudaMemsetAsync(h->d_intensity, 0, sizeof(float) * ncol * h->nDir, h->stream));
  if (h->d_intByComp && byComp)
    CUDA_OK(h, cudaMemsetAsync(h->d_intByComp, 0, sizeof(float) * ncol * h->nDir * (h->nc + 1), h->stream));
  if (h->d_excess) CUDA_OK(h, cudaMemsetAsync(h->d_excess, 0, sizeof(float) * (size_t)(h->nc + 1) * h->nDir, h->stream));
  CUDA_OK(h, cudaMemsetAsync(h->d_next, 0, sizeof(unsigned long long), h->stream));

  Problem p;
  fill_problem(h, p);
  p.src = src;
  p.key0 = key0;
  p.key1 = key1;
  p.firstPhoton = 0;
  if (p.repK > 1) {  // (copies are left zeroed by k_fold_replicas; a new or larger buffer is zeroed here)
    const size_t need = (size_t)(p.repK - 1) * p.repStride;
    if (h->repN < need) {
      dfree(h->d_rep);
      CUDA_OK(h, cudaMalloc(&h->d_rep, sizeof(float) * need));
      h->repN = need;
    }
    CUDA_OK(h, cudaMemsetAsync(h->d_rep, 0, sizeof(float) * need, h->stream));
    p.rep = h->d_rep;
  }
  auto fold_replicas = [&]() {  // after every transport launch: the copies join the tally arrays
    if (p.repK <= 1) return;
    const int K1 = p.repK - 1;
    const size_t st = (size_t)p.repStride;
    auto one = [&](float* main, size_t n, int off) {
      if (n) k_fold_replicas<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(main, h->d_rep + off, n, K1, st);
      h->otherLaunches++;
    };
    one(h->d_fluxUp, ncol, p.repOff[TAL_UP]);
    one(h->d_fluxDown, ncol, p.repOff[TAL_DOWN]);
    if (nD) one(h->d_intensity, ncol * nD, p.repOff[TAL_INT]);
    if (h->volAbsDirty) one(h->d_volAbs, ncell, p.repOff[TAL_VOL]);
  };
  // time the transport kernel with CUDA events on this handle's stream
  if (h->evUsed + 2 > h->ev.size()) {
    cudaEvent_t a, b;
    CUDA_OK(h, cudaEventCreate(&a));
    CUDA_OK(h, cudaEventCreate(&b));
    h->ev.push_back(a);
    h->ev.push_back(b);
  }
  CUDA_OK(h, cudaEventRecord(h->ev[h->evUsed], h->stream));
  int rc = I3RC_SUCCESS;
  // The batch is traced in pieces (photon ids [off, off + len), one launch each) for two reasons:
  //  * float32 tallies: an element that receives more than ~2^24 increments stops growing (a 1-column plane-parallel
  //    domain with 1e8 photons), and small increments are lost long before that.  Pieces are kept below 2^20 photons per
  //    column and folded into float64 sums in between;
  //  * hand-filled photon arrays: a first twentieth is copied and traced while the next quarter is being copied, and
  //    that while the rest is (every piece's copy runs under the kernel of the piece before: 0.6 ms of the 13 ms that
  //    320 MB take over PCIe stay exposed).
  const long long n = src.n;
  // (a spotlight or an internal source puts all photons into a few columns whatever the size of the domain)
  const bool concentrated = src.kind == I3RC_SRC_SPOTLIGHT || src.kind == I3RC_SRC_INTERNAL_FLUX || src.kind == I3RC_SRC_INTERNAL_INTENSITY;
  const long long perPiece = concentrated ? (1 << 22) : std::max<long long>(1 << 19, (long long)ncol << 20);
  const bool arrays = src.kind == I3RC_SRC_ARRAYS;
  std::vector<long long> cuts{0};
  if (arrays && n >= (1 << 20)) {
    // (2 %, 8 %, 30 % of the batch, then the rest: the first piece only has to feed the photon slots of the grid)
    for (const long long per1000 : {20LL, 80LL, 300LL}) {
      const long long cut = ((n * per1000 / 1000 + 127) / 128) * 128;
      if (cut > cuts.back() && cut < n && cut - cuts.back() <= perPiece) cuts.push_back(cut);
    }
  }
  while (n - cuts.back() > perPiece) cuts.push_back(cuts.back() + perPiece);
  cuts.push_back(n);
  const int nPieces = (int)cuts.size() - 1;
  const bool fold = n > perPiece;  // some element may see too many increments for float32
  struct Tally {
    float* f;
    size_t n;
  };
  std::vector<Tally> tallies{{h->d_fluxUp, ncol}, {h->d_fluxDown, ncol}, {h->d_fluxAbs, ncol}, {h->d_volAbs, ncell}};
  if (h->d_intensity) tallies.push_back({h->d_intensity, ncol * h->nDir});
  if (h->d_intByComp && byComp) tallies.push_back({h->d_intByComp, ncol * h->nDir * (h->nc + 1)});
  if (h->d_excess) tallies.push_back({h->d_excess, (size_t)(h->nc + 1) * h->nDir});
  size_t nTally = 0;
  for (auto& t : tallies) nTally += t.n;
  if (fold) {
    if (h->foldN < nTally) {
      dfree(h->d_fold);
      CUDA_OK(h, cudaMalloc(&h->d_fold, sizeof(double) * nTally));
      h->foldN = nTally;
    }
    CUDA_OK(h, cudaMemsetAsync(h->d_fold, 0, sizeof(double) * nTally, h->stream));
  }
  // Hand-filled arrays of a batch that needs no folding are traced by ONE launch that starts when the first piece has
  // arrived and is told, by the copy engine, how far the upload has got (SourceDev::avail): no kernel ever waits for
  // another kernel, only for the DMA engine, which does not need an SM.
  const bool streamed = arrays && !fold && nPieces > 1;
  if (arrays) {
    CUDA_OK(h, cudaEventRecord(h->computeDone, h->stream));  // the previous batch may still be reading the arrays
    CUDA_OK(h, cudaStreamWaitEvent(h->copyStream, h->computeDone, 0));
    // one copy per piece, back to back on the copy stream, each followed by an event its kernel waits for
    while ((int)h->copyDone.size() < nPieces) {
      cudaEvent_t e;
      CUDA_OK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->copyDone.push_back(e);
    }
    if (streamed) {
      if (!h->d_avail) {
        CUDA_OK(h, cudaMalloc(&h->d_avail, sizeof(unsigned long long)));
        CUDA_OK(h, cudaHostAlloc(&h->h_avail, sizeof(unsigned long long) * 64, cudaHostAllocDefault));
      }
      CUDA_OK(h, cudaMemsetAsync(h->d_avail, 0, sizeof(unsigned long long), h->copyStream));
    }
    for (int c = 0; c < nPieces; c++) {
      const long long len = cuts[c + 1] - cuts[c];
      for (int k = 0; k < 5 && len > 0; k++)
        CUDA_OK(h, cudaMemcpyAsync(h->d_srcArrays + k * n + cuts[c], h->hostArrays[k] + cuts[c], sizeof(float) * len,
                                   cudaMemcpyHostToDevice, h->copyStream));
      if (streamed && c < 64) {
        h->h_avail[c] = (unsigned long long)cuts[c + 1];
        CUDA_OK(h, cudaMemcpyAsync(h->d_avail, h->h_avail + c, sizeof(unsigned long long), cudaMemcpyHostToDevice, h->copyStream));
      }
      CUDA_OK(h, cudaEventRecord(h->copyDone[c], h->copyStream));
    }
  }
  if (streamed) {  // one launch over the whole batch
    CUDA_OK(h, cudaStreamWaitEvent(h->stream, h->copyDone[0], 0));
    p.src.n = n;
    p.src.avail = h->d_avail;
    p.firstPhoton = 0;
    rc = launch_transport(h, p);
    h->traceLaunches++;
    fold_replicas();
    if (rc == I3RC_SUCCESS && h->volAbsDirty) {
      k_abs_from_volume<<<(unsigned)((ncol + 127) / 128), 128, 0, h->stream>>>(h->nz, ncol, h->d_volAbs, h->d_fluxAbs);
      h->otherLaunches++;
    }
  }
  for (int c = 0; c < nPieces && rc == I3RC_SUCCESS && !streamed; c++) {
    const long long off = cuts[c], len = cuts[c + 1] - cuts[c];
    if (len <= 0) continue;
    if (arrays) CUDA_OK(h, cudaStreamWaitEvent(h->stream, h->copyDone[c], 0));
    if (c) CUDA_OK(h, cudaMemsetAsync(h->d_next, 0, sizeof(unsigned long long), h->stream));
    p.src.n = len;
    if (arrays) {
      p.src.ax = src.ax + off;
      p.src.ay = src.ay + off;
      p.src.az = src.az + off;
      p.src.amu = src.amu + off;
      p.src.aphi = src.aphi + off;
    }
    p.firstPhoton = off;
    rc = launch_transport(h, p);
    h->traceLaunches++;
    fold_replicas();
    if (rc == I3RC_SUCCESS && h->volAbsDirty) {  // absorbed flux per column = the column sums of the volume absorption
      k_abs_from_volume<<<(unsigned)((ncol + 127) / 128), 128, 0, h->stream>>>(h->nz, ncol, h->d_volAbs, h->d_fluxAbs);
      h->otherLaunches++;
    }
    if (fold && rc == I3RC_SUCCESS) {
      size_t o = 0;
      for (auto& t : tallies) {
        k_fold_tally<<<(unsigned)((t.n + 255) / 256), 256, 0, h->stream>>>(t.f, h->d_fold + o, t.n, c == nPieces - 1);
        o += t.n;
        h->otherLaunches++;
      }
    }
  }
  if (rc != I3RC_SUCCESS) return rc;
  CUDA_OK(h, cudaGetLastError());
  CUDA_OK(h, cudaEventRecord(h->ev[h->evUsed + 1], h->stream));
  h->evUsed += 2;

  if (nD && h->limitContrib) {  // MCRT:327-347
    int rows = (h->nc + 1) * nD;
    if (ensure_scratch(h, rows) != I3RC_SUCCESS) return I3RC_FAILURE;
    k_slab_sums<<<rows, 256, 0, h->stream>>>(h->d_intByComp, ncol, h->d_scratch);
    dim3 g((unsigned)((ncol + 255) / 256), nD);
    k_redistribute_excess<<<g, 256, 0, h->stream>>>(nD, h->nc + 1, ncol, h->d_excess, h->d_scratch, h->d_intensity,
                                                    h->d_intByComp);
    h->otherLaunches += 2;
  }
  NormArgs a;
  a.nx = h->nx;
  a.ny = h->ny;
  a.nz = h->volAbsDirty ? h->nz : 0;  // (an all-zero volume absorption needs no normalisation)
  a.nDir = nD;
  a.nc = h->nc;
  a.xyRegular = h->xyRegular;
  a.trackByComponent = byComp && nD;
  a.numPhotons = (float)src.n;
  a.xe = h->d_xe;
  a.ye = h->d_ye;
  a.ze = h->d_ze;
  a.fluxUp = h->d_fluxUp;
  a.fluxDown = h->d_fluxDown;
  a.fluxAbs = h->d_fluxAbs;
  a.volAbs = h->d_volAbs;
  a.intensity = h->d_intensity;
  a.intByComp = h->d_intByComp;
  k_normalize<<<(unsigned)((ncol + 127) / 128), 128, 0, h->stream>>>(a);
  h->otherLaunches++;
  CUDA_OK(h, cudaGetLastError());
  return I3RC_SUCCESS;
}

int prepare_compute(i3rc_integrator* h, const i3rc_photon_source* src, SourceDev& sd) {
  if (!h || !h->readyToCompute) {
    if (h) h->message = "computeRadiativeTransfer: problem not completely specified.";
    return I3RC_FAILURE;
  }
  CUDA_OK(h, cudaSetDevice(h->device));
  int rc = validate_source(h, src);
  if (rc != I3RC_SUCCESS) return rc;
  if (!h->useRayTracing && !(h->maxExt > 0.0f))
    return fail(h, "computeRadiativeTransfer: maximum cross-section needs a domain with extinction > 0.");
  rc = tabulate(h);
  if (rc != I3RC_SUCCESS) return rc;
  // lower bounds of the optical path to the top, per radiance direction and cell (Problem::leLB): regular grids, roulette
  // for intensity; at most 8 GB (twelve slanted directions on 512x512x256 are 3.2 GB)
  // (few-column domains -- step cloud, radar cloud -- have rays of a handful of crossings: tracing them is as cheap as
  //  asking, measured -2.5 %; `le_lower_bound` = 2 forces the bounds there too)
  if (!h->leLBValid && h->leLowerBound && h->computeIntensity && h->useRRIntensity && h->xyRegular && h->zRegular && h->nDir > 0 &&
      ((size_t)h->nx * h->ny >= 4096 || h->leLowerBound == 2)) {
    const size_t ncell = (size_t)h->nx * h->ny * h->nz;
    if (ncell * h->nDir * sizeof(float) <= ((size_t)8 << 30)) {
      dfree(h->d_leLB);
      dfree(h->d_leUB);
      CUDA_OK(h, cudaMalloc(&h->d_leLB, sizeof(float) * ncell * h->nDir));
      const unsigned g = (unsigned)((ncell + 127) / 128);
      // (the full depth of the domain where that is cheap: 2 M cells x 3 directions take ~10 ms -- then the upper bound
      //  exists too --; LE_LB_LAYERS layers otherwise)
      const bool full = ncell * h->nDir <= ((size_t)64 << 20);
      if (full) CUDA_OK(h, cudaMalloc(&h->d_leUB, sizeof(float) * ncell * h->nDir));
      k_le_path_bounds<<<g, 128, 0, h->stream>>>(h->nx, h->ny, h->nz, h->deltaX, h->deltaY, h->deltaZ, h->d_ext, h->d_dirs, h->nDir,
                                                 full ? h->nz : LE_LB_LAYERS, h->d_leLB, h->d_leUB);
      h->otherLaunches++;
      CUDA_OK(h, cudaGetLastError());
      h->leLBValid = true;
    }
  }
  return fill_source(h, src, sd);
}

int fetch_counters(i3rc_integrator* h) {
  unsigned long long c[CNT_N];
  CUDA_OK(h, cudaMemcpyAsync(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  i3rc_counters& o = h->counters;
  o.photons = c[CNT_PHOTONS];
  o.bad = c[CNT_BAD];
  o.crossings_photon = c[CNT_CROSS_PH];
  o.crossings_intensity = c[CNT_CROSS_LE];
  o.collisions = c[CNT_COLL];
  o.absorptions = c[CNT_ABS];
  o.contributions = c[CNT_CONTRIB];
  o.exits_top = c[CNT_TOP];
  o.surface_hits = c[CNT_SURF];
  o.rng_draws = c[CNT_RNG];
  o.roulette_kills = c[CNT_KILL];
  o.null_collisions = c[CNT_NULL];
  o.cells_skipped = c[CNT_SKIP];
  o.cells_skipped_intensity = c[CNT_SKIP_LE];
#ifdef I3RC_SCHED_STATS
  {
    const double rounds = (double)c[ST_ROUNDS], pairs = (double)c[ST_PAIRS], batch = (double)c[ST_BATCH];
    fprintf(stderr,
            "[sched] rounds %.4g  lanes running at round start %.2f  step pairs per round %.2f  lanes per step pair %.2f\n"
            "[sched] idle lanes left without a task per round %.2f  tasks left in the ring per round %.2f\n"
            "[sched] event batches %.4g (%.3f per round)  events per batch %.2f  alive after the event %.2f  "
            "local-estimate tasks per batch %.2f  suspensions per batch %.3f  births per batch with births %.2f (%.3f of the batches)\n",
            rounds, c[ST_START_RUN] / rounds, pairs / rounds, c[ST_LANE_PAIRS] / pairs, c[ST_STARVED] / rounds,
            c[ST_RING_LEFT] / rounds, batch, batch / rounds, c[ST_BATCH_HAS] / batch, c[ST_BATCH_ALIVE] / batch,
            c[ST_LE_PUSH] / batch, c[ST_SUSP] / batch, c[ST_BIRTHS] / (double)(c[ST_BIRTH_BATCH] ? c[ST_BIRTH_BATCH] : 1),
            c[ST_BIRTH_BATCH] / batch);
  }
#endif
  return I3RC_SUCCESS;
}

void seed_to_key(const int32_t* seed, int nseed, uint32_t& k0, uint32_t& k1) {
  k0 = nseed > 0 ? (uint32_t)seed[0] : 0u;
  k1 = nseed > 1 ? (uint32_t)seed[1] : 0x9E3779B9u;
  for (int i = 2; i < nseed; i++) {  // longer seed vectors are folded in
    k0 = k0 * 0x01000193u ^ (uint32_t)seed[i];
    k1 = (k1 << 5 | k1 >> 27) ^ (uint32_t)seed[i] * 0x85EBCA6Bu;
  }
}

// packed layout of the batch-moment buffer (doubles); every block is [sum(x) | sum(x*x)]
struct StatsLayout {
  size_t mean[3], flux[3], prof, rad, mrad, vol, total;
};
StatsLayout stats_layout(const i3rc_integrator* h, int nD, bool withVolume) {
  StatsLayout L;
  size_t ncol = (size_t)h->nx * h->ny, o = 0;
  for (int i = 0; i < 3; i++) {
    L.mean[i] = o;
    o += 2;
  }
  for (int i = 0; i < 3; i++) {
    L.flux[i] = o;
    o += 2 * ncol;
  }
  L.prof = o;
  o += 2 * (size_t)h->nz;
  L.rad = o;
  o += 2 * ncol * nD;
  L.mrad = o;
  o += 2 * (size_t)nD;
  L.vol = o;
  if (withVolume) o += 2 * ncol * h->nz;
  L.total = o;
  return L;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* i3rc_version(void) { return "i3rc_b200 0.1.0 (sm_100a)"; }

int i3rc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
int i3rc_set_device(int device) { return cudaSetDevice(device) == cudaSuccess ? I3RC_SUCCESS : I3RC_FAILURE; }

const char* i3rc_last_message(const i3rc_integrator* h) { return h ? h->message.c_str() : g_message.c_str(); }

int i3rc_new_Integrator(int nx, int ny, int nz, int nc, const float* xPos, const float* yPos, const float* zPos,
                        const float* totalExt, const float* cumExt, const float* ssa, const int32_t* pfIndex,
                        i3rc_integrator** out) {
  if (out) *out = nullptr;
  if (!have_device()) {
    g_message = "new_Integrator: no CUDA device (the B200 integrator has no CPU fallback)";
    return I3RC_FAILURE;
  }
  if (!out || nx < 1 || ny < 1 || nz < 1 || nc < 1 || !xPos || !yPos || !zPos || !totalExt || !cumExt || !ssa || !pfIndex ||
      !valid_edges(xPos, nx) || !valid_edges(yPos, ny) || !valid_edges(zPos, nz)) {
    g_message = "new_Integrator: Problems reading domain.";
    return I3RC_FAILURE;
  }
  size_t ncell = (size_t)nx * ny * nz;
  std::vector<int> maxPf(nc, 0);
  bool absorbs = false;
  for (int c = 0; c < nc; c++) {  // (the dense form has cumulative fractions instead of extinctions: checked against totalExt)
    const char* bad = validate_optical_arrays(totalExt, ssa + c * ncell, pfIndex + c * ncell, ncell, &maxPf[c], &absorbs);
    if (bad) {
      g_message = bad;
      return I3RC_FAILURE;
    }
  }
  i3rc_integrator* h = make_handle();
  h->nx = nx;
  h->ny = ny;
  h->nz = nz;
  h->nc = nc;
  h->maxPfIndex = maxPf;
  h->absorbing = absorbs;
  h->absorbing = absorbs;
  h->xe.assign(xPos, xPos + nx + 1);
  h->ye.assign(yPos, yPos + ny + 1);
  h->ze.assign(zPos, zPos + nz + 1);
  int rc = I3RC_SUCCESS;
  if (upload(&h->d_ext, totalExt, ncell, h->stream) != cudaSuccess || upload(&h->d_cum, cumExt, ncell * nc, h->stream) != cudaSuccess ||
      upload(&h->d_ssa, ssa, ncell * nc, h->stream) != cudaSuccess || upload(&h->d_pf, pfIndex, ncell * nc, h->stream) != cudaSuccess)
    rc = fail(h, "new_Integrator: out of device memory");
  if (rc == I3RC_SUCCESS) rc = finish_new_integrator(h);
  if (rc == I3RC_FAILURE) {
    g_message = h->message;
    i3rc_finalize_Integrator(h);
    return rc;
  }
  *out = h;
  return rc;
}

int i3rc_new_Integrator_components(int nx, int ny, int nz, const float* xPos, const float* yPos, const float* zPos,
                                   int nc, const i3rc_component* comps, i3rc_integrator** out) {
  if (out) *out = nullptr;
  if (!have_device()) {
    g_message = "new_Integrator: no CUDA device (the B200 integrator has no CPU fallback)";
    return I3RC_FAILURE;
  }
  if (!out || nx < 1 || ny < 1 || nz < 1 || nc < 1 || !comps || !xPos || !yPos || !zPos || !valid_edges(xPos, nx) ||
      !valid_edges(yPos, ny) || !valid_edges(zPos, nz)) {
    g_message = "new_Integrator: Problems reading domain.";
    return I3RC_FAILURE;
  }
  for (int c = 0; c < nc; c++)
    if (comps[c].z_level_base < 1 || comps[c].z_level_base + comps[c].nz - 1 > nz || !comps[c].extinction ||
        !comps[c].ssa || !comps[c].phase_index) {
      g_message = "getOpticalPropertiesByComponent: component does not conform to the domain.";
      return I3RC_FAILURE;
    }
  std::vector<int> maxPf(nc, 0);
  bool absorbs = false;
  for (int c = 0; c < nc; c++) {
    const size_t n = (comps[c].horizontally_uniform ? 1 : (size_t)nx * ny) * (size_t)comps[c].nz;
    const char* bad = validate_optical_arrays(comps[c].extinction, comps[c].ssa, comps[c].phase_index, n, &maxPf[c], &absorbs);
    if (!bad && maxPf[c] > comps[c].table.n_entries) bad = "validateOpticalComponent: phase function index is out of bounds";
    if (bad) {
      g_message = bad;
      return I3RC_FAILURE;
    }
  }
  i3rc_integrator* h = make_handle();
  h->nx = nx;
  h->ny = ny;
  h->nz = nz;
  h->nc = nc;
  h->maxPfIndex = maxPf;
  h->absorbing = absorbs;
  h->xe.assign(xPos, xPos + nx + 1);
  h->ye.assign(yPos, yPos + ny + 1);
  h->ze.assign(zPos, zPos + nz + 1);
  size_t ncell = (size_t)nx * ny * nz, ncol = (size_t)nx * ny;
  int rc = I3RC_SUCCESS;
  std::vector<ComponentDev> cd(nc);
  std::vector<void*> temps;
  auto body = [&]() -> int {
    CUDA_OK(h, cudaMalloc(&h->d_ext, sizeof(float) * ncell));
    CUDA_OK(h, cudaMalloc(&h->d_cum, sizeof(float) * ncell * nc));
    CUDA_OK(h, cudaMalloc(&h->d_ssa, sizeof(float) * ncell * nc));
    CUDA_OK(h, cudaMalloc(&h->d_pf, sizeof(int) * ncell * nc));
    for (int c = 0; c < nc; c++) {
      size_t n = (comps[c].horizontally_uniform ? 1 : ncol) * (size_t)comps[c].nz;
      float *e = nullptr, *s = nullptr;
      int* f = nullptr;
      CUDA_OK(h, upload(&e, comps[c].extinction, n, h->stream));
      temps.push_back(e);
      CUDA_OK(h, upload(&s, comps[c].ssa, n, h->stream));
      temps.push_back(s);
      CUDA_OK(h, upload(&f, (const int*)comps[c].phase_index, n, h->stream));
      temps.push_back(f);
      cd[c] = ComponentDev{e, s, f, comps[c].horizontally_uniform, comps[c].z_level_base - 1, comps[c].nz};
    }
    ComponentDev* d_cd = nullptr;
    CUDA_OK(h, upload(&d_cd, cd.data(), cd.size(), h->stream));
    temps.push_back(d_cd);
    k_expand_components<<<(unsigned)((ncell + 255) / 256), 256, 0, h->stream>>>(nx, ny, nz, nc, d_cd, h->d_ext, h->d_cum,
                                                                               h->d_ssa, h->d_pf);
    h->otherLaunches++;
    CUDA_OK(h, cudaGetLastError());
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return I3RC_SUCCESS;
  };
  rc = body();
  for (void* t : temps) cudaFree(t);
  if (rc == I3RC_SUCCESS) rc = finish_new_integrator(h);
  if (rc == I3RC_SUCCESS)
    for (int c = 0; c < nc && rc != I3RC_FAILURE; c++) rc = i3rc_set_phase_table(h, c, &comps[c].table);
  if (rc == I3RC_FAILURE) {
    g_message = h->message;
    i3rc_finalize_Integrator(h);
    return rc;
  }
  *out = h;
  return rc;
}

int i3rc_set_phase_table(i3rc_integrator* h, int comp, const i3rc_phase_table* t) {
  if (!h) return I3RC_FAILURE;
  if (comp < 0 || comp >= h->nc || !t) return fail(h, "set_phase_table: no such component");
  if (t->n_entries < 1) return fail(h, "set_phase_table: phase function table is not ready.");
  if (t->n_entries > 65535) return fail(h, "set_phase_table: at most 65535 entries per phase function table in this implementation.");
  if (comp < (int)h->maxPfIndex.size() && h->maxPfIndex[comp] > t->n_entries)
    return fail(h, "validateOpticalComponent: phase function index is out of bounds");
  CUDA_OK(h, cudaSetDevice(h->device));
  HostTable& o = h->tables[comp];
  free_table(o);
  o = HostTable();
  o.kind = t->kind;
  o.nEntries = t->n_entries;
  o.nodesPerEntry.resize(t->n_entries);
  if (t->kind == 1) {
    if (!t->coef_offsets) return fail(h, "set_phase_table: Legendre table without offsets");
    o.offsets.assign(t->coef_offsets, t->coef_offsets + t->n_entries + 1);
    int total = o.offsets.back();
    if (total > 0 && !t->coefs) return fail(h, "set_phase_table: Legendre table without coefficients");
    o.coefs.assign(t->coefs, t->coefs + total);
    if (o.coefs.empty()) o.coefs.push_back(0.0f);
    for (int e = 0; e < t->n_entries; e++) {
      int nm = o.offsets[e + 1] - o.offsets[e];
      if (nm < 0) return fail(h, "set_phase_table: offsets must be increasing");
      if (nm > 1 && (o.coefs[o.offsets[e]] > 1.0f || o.coefs[o.offsets[e]] < -1.0f))
        return fail(h, "newPhaseFunction: Asymmetery parameter out of bounds.");
      o.nodesPerEntry[e] = std::max(nm, 2);
      o.maxNodes = std::max(o.maxNodes, o.nodesPerEntry[e]);
    }
    CUDA_OK(h, upload(&o.d_offsets, o.offsets.data(), o.offsets.size(), h->stream));
    CUDA_OK(h, upload(&o.d_coefs, o.coefs.data(), o.coefs.size(), h->stream));
  } else if (t->kind == 2) {
    int n = t->n_angles;
    if (n < 2 || !t->angles || !t->values) return fail(h, "newPhaseFunctionTable: Number of scattering angles and phase function values must match.");
    const float pi = 3.141592654f;
    for (int i = 0; i < n; i++)
      if (t->angles[i] < 0.0f || t->angles[i] > pi) return fail(h, "newPhaseFunctionTable: ScatteringAngle out of bounds.");
    if (fabsf(t->angles[0]) > h_spacing(0.0f)) return fail(h, "newPhaseFunctionTable: First scattering angle must be min value");
    if (fabsf(t->angles[n - 1] - pi) > h_spacing(pi)) return fail(h, "newPhaseFunctionTable: Last scattering angle must be max value");
    for (int i = 0; i + 1 < n; i++)
      if (!(t->angles[i + 1] - t->angles[i] > 0.0f)) return fail(h, "newPhaseFunctionTable: Scattering angle must be increasing, unique.");
    for (size_t i = 0; i < (size_t)n * t->n_entries; i++)
      if (t->values[i] < 0.0f) return fail(h, "newPhaseFunctionTable: Negative phase function values supplied.");
    o.nAngles = n;
    o.angles.assign(t->angles, t->angles + n);
    o.values.assign(t->values, t->values + (size_t)n * t->n_entries);
    for (int e = 0; e < t->n_entries; e++) o.nodesPerEntry[e] = n;
    o.maxNodes = n;
    CUDA_OK(h, upload(&o.d_angles, o.angles.data(), o.angles.size(), h->stream));
    CUDA_OK(h, upload(&o.d_values, o.values.data(), o.values.size(), h->stream));
    // normalised by the constructor and again by the copy the integrator keeps
    // (scatteringPhaseFunctions.f95:322-325, 433; opticalProperties.f95:525)
    CUDA_OK(h, normalize_tabulated(o.d_angles, o.d_values, n, t->n_entries, 2, h->stream));
    h->otherLaunches += 2;
  } else {
    return fail(h, "set_phase_table: unknown table kind");
  }
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  free_matrix(h->inv[comp]);
  free_matrix(h->fwd[comp]);
  free_matrix(h->fwdOrig[comp]);
  h->tableDescDirty = true;
  h->tableDescDirtyForMax = true;
  h->message.clear();
  return I3RC_SUCCESS;
}

void i3rc_finalize_Integrator(i3rc_integrator* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  i3rc_comm_finalize(h);
  dfree(h->d_xe);
  dfree(h->d_ye);
  dfree(h->d_ze);
  dfree(h->d_ext);
  dfree(h->d_extRaw);
  dfree(h->d_leLB);
  dfree(h->d_leUB);
  dfree(h->d_colTau);
  dfree(h->d_extJ);
  dfree(h->d_zslab);
  dfree(h->d_extZ);
  dfree(h->d_zlut);
  dfree(h->d_cum);
  dfree(h->d_ssa);
  dfree(h->d_pf);
  dfree(h->d_surf_x);
  dfree(h->d_surf_y);
  dfree(h->d_surf_p);
  for (auto& t : h->tables) free_table(t);
  for (auto& m : h->inv) free_matrix(m);
  for (auto& m : h->fwd) free_matrix(m);
  for (auto& m : h->fwdOrig) free_matrix(m);
  dfree(h->d_tableDesc);
  dfree(h->d_dirs);
  dfree(h->d_fluxUp);
  dfree(h->d_fluxDown);
  dfree(h->d_fluxAbs);
  dfree(h->d_volAbs);
  dfree(h->d_intensity);
  dfree(h->d_intByComp);
  dfree(h->d_excess);
  dfree(h->d_counters);
  dfree(h->d_fold);
  if (h->copyStream) {
    cudaStreamDestroy(h->copyStream);
    for (auto& e : h->copyDone) cudaEventDestroy(e);
    cudaEventDestroy(h->computeDone);
    h->copyStream = nullptr;
  }
  dfree(h->d_next);
  dfree(h->d_susp);
  dfree(h->d_rep);
  dfree(h->d_avail);
  if (h->h_avail) cudaFreeHost(h->h_avail);
  h->h_avail = nullptr;
  dfree(h->d_scratch);
  dfree(h->d_fscratch);
  dfree(h->d_srcArrays);
  dfree(h->d_stats);
  for (auto e : h->ev) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int i3rc_isReady_Integrator(const i3rc_integrator* h) { return h && h->readyToCompute; }

int i3rc_specifyParameters(i3rc_integrator* h, const i3rc_params* p) {
  if (!h || !p) return I3RC_FAILURE;
  const uint32_t m = p->present;
  auto has = [m](uint32_t b) { return (m & b) != 0; };
  bool warn = false, failed = false;
  auto W = [&](const char* s) {
    h->message = s;
    warn = true;
  };
  auto F = [&](const char* s) {
    h->message = s;
    failed = true;
  };
  // MCRT:872-947
  if (has(I3RC_P_surfaceBDRF) && has(I3RC_P_surfaceAlbedo)) F("specifyParameters: only one surface specification can be provided");
  if (has(I3RC_P_surfaceAlbedo) && (p->surfaceAlbedo > 1.0f || p->surfaceAlbedo < 0.0f)) F("specifyParameters: surface albedo out of range.");
  if (has(I3RC_P_surfaceBDRF) && !(p->surf_x && p->surf_y && p->surf_params && p->surf_nx > 0 && p->surf_ny > 0))
    F("specifyParameters: surface description isn't valid.");
  if (has(I3RC_P_minForwardTableSize) && p->minForwardTableSize < 9001)
    W("specifyParameters: minForwardTableSize less than default. Value ignored.");
  if (has(I3RC_P_minInverseTableSize) && p->minInverseTableSize < 9001)
    W("specifyParameters: minInverseTableSize less than default. Value ignored.");
  if (has(I3RC_P_hybridPhaseFunWidth) && (p->hybridPhaseFunWidth > 30.0f || p->hybridPhaseFunWidth < 0.0f))
    W("specifyParameters: hybridPhaseFunWidth out of range (0 to 30degrees).Using default (7)");
  if (has(I3RC_P_numOrdersOrigPhaseFunIntenCalcs) && p->numOrdersOrigPhaseFunIntenCalcs < 0)
    W("specifyParameters: numOrdersExactPhaseFunIntenCalcs less than 0.Using default (0)");
  if (has(I3RC_P_maxIntensityContribution) && p->maxIntensityContribution <= 0.0f)
    W("specifyParameters: maxIntensityContribution <= 0. Value is unchanged.");
  if (has(I3RC_P_intensityMus) != has(I3RC_P_intensityPhis))
    F("specifyParameters: Both or neither of intensityMus and intensityPhis must be supplied");
  if (has(I3RC_P_intensityMus) && !failed) {
    if (p->numIntensityDirections < 1 || p->numIntensityDirections > MAX_DIRS || !p->intensityMus || !p->intensityPhis) {
      F("specifyParameters: between 1 and 32 intensity directions are supported");
    } else {
      for (int i = 0; i < p->numIntensityDirections; i++) {
        if (p->intensityMus[i] < -1.0f || p->intensityMus[i] > 1.0f) F("specifyParameters: intensityMus must be between -1 and 1");
        if (fabsf(p->intensityMus[i]) < F_TINY) F("specifyParameters: intensityMus can't be 0 (directly sideways)");
        if (p->intensityPhis[i] < 0.0f || p->intensityPhis[i] > 360.0f) F("specifyParameters: intensityPhis must be between 0 and 360");
      }
    }
  }
  if (has(I3RC_P_computeIntensity)) {
    if (!p->computeIntensity && has(I3RC_P_intensityMus))
      W("specifyParameters: intensity directions *and* computeIntensity set to false.Will compute intensity at given angles.");
    if (p->computeIntensity && !has(I3RC_P_intensityMus) && h->dirs.empty())
      F("specifyParameters: Can't compute intensity without specifying directions.");
  }
  if (failed) return I3RC_FAILURE;
  CUDA_OK(h, cudaSetDevice(h->device));

  // MCRT:950-1067
  if (has(I3RC_P_surfaceAlbedo)) {
    h->surfaceAlbedo = p->surfaceAlbedo;
    h->useSurfaceBDRF = false;
  } else if (has(I3RC_P_surfaceBDRF)) {
    h->surf_nx = p->surf_nx;
    h->surf_ny = p->surf_ny;
    h->surf_x.assign(p->surf_x, p->surf_x + p->surf_nx + 1);
    h->surf_y.assign(p->surf_y, p->surf_y + p->surf_ny + 1);
    h->surf_p.assign(p->surf_params, p->surf_params + (size_t)p->surf_nx * p->surf_ny);
    CUDA_OK(h, upload(&h->d_surf_x, h->surf_x.data(), h->surf_x.size(), h->stream));
    CUDA_OK(h, upload(&h->d_surf_y, h->surf_y.data(), h->surf_y.size(), h->stream));
    CUDA_OK(h, upload(&h->d_surf_p, h->surf_p.data(), h->surf_p.size(), h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    h->useSurfaceBDRF = true;
  }
  if (has(I3RC_P_useRayTracing)) h->useRayTracing = p->useRayTracing != 0;
  if (has(I3RC_P_minForwardTableSize)) h->minForwardTableSize = std::max(p->minForwardTableSize, 9001);
  if (has(I3RC_P_minInverseTableSize)) h->minInverseTableSize = std::max(p->minInverseTableSize, 9001);
  if (has(I3RC_P_useRussianRoulette)) h->useRussianRoulette = p->useRussianRoulette != 0;
  if (has(I3RC_P_useRussianRouletteForIntensity)) h->useRRIntensity = p->useRussianRouletteForIntensity != 0;
  if (has(I3RC_P_zetaMin)) {
    if (p->zetaMin < 0.0f) {
      W("specifyParameters: zetaMin must be >= 0. Value is unchanged.");
    } else {
      h->zetaMin = p->zetaMin;
      if (p->zetaMin > 1.0f) W("specifyParameters: zetaMin > 1. That's kind of large.");
    }
  }
  if (has(I3RC_P_useHybridPhaseFunsForIntenCalcs)) {
    bool v = p->useHybridPhaseFunsForIntenCalcs != 0;
    if (v != h->useHybrid)  // the forward tables depend on it: re-tabulate
      for (auto& mtx : h->fwd) free_matrix(mtx);
    h->useHybrid = v;
  }
  if (has(I3RC_P_hybridPhaseFunWidth)) {
    if (p->hybridPhaseFunWidth > 0.0f && p->hybridPhaseFunWidth < 30.0f)
      h->hybridWidth = p->hybridPhaseFunWidth;
    else
      h->hybridWidth = 7.0f;
    for (auto& mtx : h->fwd) free_matrix(mtx);  // MCRT:996-1004: re-tabulate
    h->tableDescDirty = true;
  h->tableDescDirtyForMax = true;
  }
  if (has(I3RC_P_numOrdersOrigPhaseFunIntenCalcs))
    h->numOrdersOrig = p->numOrdersOrigPhaseFunIntenCalcs >= 0 ? p->numOrdersOrigPhaseFunIntenCalcs : 0;
  if (has(I3RC_P_limitIntensityContributions)) h->limitContrib = p->limitIntensityContributions != 0;
  if (has(I3RC_P_maxIntensityContribution) && p->maxIntensityContribution > 0.0f) h->maxContrib = p->maxIntensityContribution;
  size_t ncol = (size_t)h->nx * h->ny;
  if (has(I3RC_P_intensityMus)) {  // MCRT:1026-1045
    h->nDir = p->numIntensityDirections;
    h->dirs.assign((size_t)h->nDir * DIR_STRIDE, 0.0f);
    for (int i = 0; i < h->nDir; i++) {
      float mu = p->intensityMus[i], phi = p->intensityPhis[i] * F_PI / 180.0f;
      float st = sqrtf(1.0f - mu * mu);
      float* d = &h->dirs[(size_t)i * DIR_STRIDE];
      d[0] = st * cosf(phi);
      d[1] = st * sinf(phi);
      d[2] = mu;
      for (int a = 0; a < 3; a++) d[3 + a] = fabsf(d[a]) >= 2.0f * F_TINY ? 1.0f / fabsf(d[a]) : INFINITY;
      d[6] = 4.0f * F_PI * fabsf(mu);
      d[7] = 1.0f / d[6];
    }
    CUDA_OK(h, upload(&h->d_dirs, h->dirs.data(), h->dirs.size(), h->stream));
    h->leLBValid = false;
    dfree(h->d_intensity);
    dfree(h->d_intByComp);
    CUDA_OK(h, cudaMalloc(&h->d_intensity, sizeof(float) * ncol * h->nDir));
    CUDA_OK(h, cudaMalloc(&h->d_intByComp, sizeof(float) * ncol * h->nDir * (h->nc + 1)));
    CUDA_OK(h, cudaMemsetAsync(h->d_intensity, 0, sizeof(float) * ncol * h->nDir, h->stream));
    CUDA_OK(h, cudaMemsetAsync(h->d_intByComp, 0, sizeof(float) * ncol * h->nDir * (h->nc + 1), h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    h->computeIntensity = true;
  }
  if (has(I3RC_P_computeIntensity) && !p->computeIntensity && !has(I3RC_P_intensityMus)) {  // MCRT:1052-1057
    h->dirs.clear();
    h->nDir = 0;
    dfree(h->d_dirs);
    dfree(h->d_intensity);
    dfree(h->d_intByComp);
    h->computeIntensity = false;
  }
  if (h->computeIntensity && h->limitContrib) {  // MCRT:1060-1064
    dfree(h->d_excess);
    CUDA_OK(h, cudaMalloc(&h->d_excess, sizeof(float) * (size_t)(h->nc + 1) * h->nDir));
  }
  if (!warn) h->message.clear();
  return warn ? I3RC_WARNING : I3RC_SUCCESS;
}

int i3rc_tabulate(i3rc_integrator* h) {
  if (!h) return I3RC_FAILURE;
  CUDA_OK(h, cudaSetDevice(h->device));
  return tabulate(h);
}

int i3rc_get_table(i3rc_integrator* h, int which, int comp, float* out, int* nSteps, int* nEntries) {
  if (!h || comp < 0 || comp >= h->nc) return I3RC_FAILURE;
  DevMatrix& m = which == 0 ? h->inv[comp] : which == 1 ? h->fwd[comp] : h->fwdOrig[comp];
  if (!m.d) return fail(h, "get_table: table has not been tabulated");
  if (nSteps) *nSteps = m.nSteps;
  if (nEntries) *nEntries = m.nEntries;
  if (out) {
    CUDA_OK(h, cudaMemcpyAsync(out, m.d, sizeof(float) * (size_t)m.nSteps * m.nEntries, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
  }
  return I3RC_SUCCESS;
}

int i3rc_set_inverse_table(i3rc_integrator* h, int comp, int nSteps, int nEntries, const float* values) {
  if (!h || comp < 0 || comp >= h->nc || nSteps < 2 || nEntries < 1 || !values) return I3RC_FAILURE;
  free_matrix(h->inv[comp]);
  CUDA_OK(h, upload(&h->inv[comp].d, values, (size_t)nSteps * nEntries, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  h->inv[comp].nSteps = nSteps;
  h->inv[comp].nEntries = nEntries;
  if (nSteps > h->minInverseTableSize) h->minInverseTableSize = nSteps;
  h->tableDescDirty = true;
  h->tableDescDirtyForMax = true;
  return I3RC_SUCCESS;
}
int i3rc_set_forward_table(i3rc_integrator* h, int comp, int nSteps, int nEntries, const float* tabulated,
                           const float* original) {
  if (!h || comp < 0 || comp >= h->nc || nSteps < 2 || nEntries < 1 || !tabulated) return I3RC_FAILURE;
  free_matrix(h->fwd[comp]);
  free_matrix(h->fwdOrig[comp]);
  CUDA_OK(h, upload(&h->fwd[comp].d, tabulated, (size_t)nSteps * nEntries, h->stream));
  CUDA_OK(h, upload(&h->fwdOrig[comp].d, original ? original : tabulated, (size_t)nSteps * nEntries, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  h->fwd[comp].nSteps = h->fwdOrig[comp].nSteps = nSteps;
  h->fwd[comp].nEntries = h->fwdOrig[comp].nEntries = nEntries;
  if (nSteps > h->minForwardTableSize) h->minForwardTableSize = nSteps;
  h->tableDescDirty = true;
  h->tableDescDirtyForMax = true;
  return I3RC_SUCCESS;
}

int i3rc_set_component_profile(i3rc_integrator* h, int comp, const float* extinction) {
  if (!h || !h->readyToCompute) return I3RC_FAILURE;
  if (comp < 0 || comp >= h->nc || !extinction) return fail(h, "set_component_profile: no such component.");
  for (int k = 0; k < h->nz; k++)
    if (!(extinction[k] >= 0.0f)) return fail(h, "set_component_profile: extinction must be >= 0.");
  CUDA_OK(h, cudaSetDevice(h->device));
  const size_t ncell = (size_t)h->nx * h->ny * h->nz;
  h->absorbing = true;  // (the new profile may switch an absorbing component on)
  if (!h->d_extRaw) {
    CUDA_OK(h, cudaMalloc(&h->d_extRaw, sizeof(float) * ncell * h->nc));
    k_recover_extinctions<<<(unsigned)((ncell + 255) / 256), 256, 0, h->stream>>>(ncell, h->nc, h->d_ext, h->d_cum, h->d_extRaw);
    h->otherLaunches++;
  }
  float* d_prof = nullptr;
  int rc = [&]() -> int {
    CUDA_OK(h, upload(&d_prof, extinction, (size_t)h->nz, h->stream));
    k_replace_profile<<<(unsigned)((ncell + 255) / 256), 256, 0, h->stream>>>(h->nx, h->ny, h->nz, h->nc, comp, d_prof, h->d_extRaw,
                                                                              h->d_ext, h->d_cum);
    h->otherLaunches++;
    CUDA_OK(h, cudaGetLastError());
    return build_gather_field(h);  // maximum extinction, 1 + epsilon nudge, the copy of totalExt the rays gather from
  }();
  cudaFree(d_prof);
  if (rc != I3RC_SUCCESS) return rc;
  h->message.clear();
  return I3RC_SUCCESS;
}

int i3rc_computeRadiativeTransfer(i3rc_integrator* h, const i3rc_photon_source* src, const int32_t* seed, int nseed) {
  SourceDev sd;
  int rc = prepare_compute(h, src, sd);
  if (rc != I3RC_SUCCESS) return rc;
  uint32_t k0, k1;
  seed_to_key(seed, nseed, k0, k1);
  CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * CNT_N, h->stream));
  rc = run_one_batch(h, sd, k0, k1);
  if (rc != I3RC_SUCCESS) return rc;
  rc = fetch_counters(h);
  if (rc != I3RC_SUCCESS) return rc;
  if (h->counters.photons <= 0) return fail(h, "computeRadiativeTransfer: Didn't process any photons.");
  h->message = "computeRadiativeTransfer: finished with photons";
  return I3RC_SUCCESS;
}

int i3rc_reportResults(i3rc_integrator* h, float* meanFluxUp, float* meanFluxDown, float* meanFluxAbsorbed,
                       float* fluxUp, float* fluxDown, float* fluxAbsorbed, float* absorbedProfile,
                       float* volumeAbsorption, float* meanIntensity, float* intensity) {
  if (!h || !h->readyToCompute) return I3RC_FAILURE;
  CUDA_OK(h, cudaSetDevice(h->device));
  size_t ncol = (size_t)h->nx * h->ny, ncell = ncol * h->nz;
  if ((meanIntensity || intensity) && !h->d_intensity) return fail(h, "reportResults: intensity information not available");
  int nD = h->d_intensity ? h->nDir : 0;
  int rows = 3 + h->nz + nD;
  if (ensure_scratch(h, rows) != I3RC_SUCCESS || ensure_fscratch(h, rows) != I3RC_SUCCESS) return I3RC_FAILURE;
  std::vector<float> host(rows, 0.0f);
  bool needMeans = meanFluxUp || meanFluxDown || meanFluxAbsorbed || absorbedProfile || meanIntensity;
  if (needMeans) {
    k_slab_sums<<<1, 256, 0, h->stream>>>(h->d_fluxUp, ncol, h->d_scratch + 0);
    k_slab_sums<<<1, 256, 0, h->stream>>>(h->d_fluxDown, ncol, h->d_scratch + 1);
    k_slab_sums<<<1, 256, 0, h->stream>>>(h->d_fluxAbs, ncol, h->d_scratch + 2);
    k_slab_sums<<<h->nz, 256, 0, h->stream>>>(h->d_volAbs, ncol, h->d_scratch + 3);
    if (nD) k_slab_sums<<<nD, 256, 0, h->stream>>>(h->d_intensity, ncol, h->d_scratch + 3 + h->nz);
    k_means_to_float<<<(rows + 127) / 128, 128, 0, h->stream>>>(h->d_scratch, rows, 1.0 / (double)ncol, h->d_fscratch);
    h->otherLaunches += 6;
    CUDA_OK(h, cudaMemcpyAsync(host.data(), h->d_fscratch, sizeof(float) * rows, cudaMemcpyDeviceToHost, h->stream));
  }
  if (fluxUp) CUDA_OK(h, cudaMemcpyAsync(fluxUp, h->d_fluxUp, sizeof(float) * ncol, cudaMemcpyDeviceToHost, h->stream));
  if (fluxDown) CUDA_OK(h, cudaMemcpyAsync(fluxDown, h->d_fluxDown, sizeof(float) * ncol, cudaMemcpyDeviceToHost, h->stream));
  if (fluxAbsorbed) CUDA_OK(h, cudaMemcpyAsync(fluxAbsorbed, h->d_fluxAbs, sizeof(float) * ncol, cudaMemcpyDeviceToHost, h->stream));
  if (volumeAbsorption)
    CUDA_OK(h, cudaMemcpyAsync(volumeAbsorption, h->d_volAbs, sizeof(float) * ncell, cudaMemcpyDeviceToHost, h->stream));
  if (intensity)
    CUDA_OK(h, cudaMemcpyAsync(intensity, h->d_intensity, sizeof(float) * ncol * nD, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  if (meanFluxUp) *meanFluxUp = host[0];
  if (meanFluxDown) *meanFluxDown = host[1];
  if (meanFluxAbsorbed) *meanFluxAbsorbed = host[2];
  if (absorbedProfile) memcpy(absorbedProfile, host.data() + 3, sizeof(float) * h->nz);
  if (meanIntensity) memcpy(meanIntensity, host.data() + 3 + h->nz, sizeof(float) * nD);
  h->message.clear();
  return I3RC_SUCCESS;
}

int i3rc_get_intensityByComponent(i3rc_integrator* h, float* out) {
  if (!h || !h->d_intByComp || !(h->trackByComponent || h->limitContrib))
    return h ? fail(h, "intensityByComponent is only tracked with limitIntensityContributions or tuning track_by_component=1") : I3RC_FAILURE;
  size_t n = (size_t)h->nx * h->ny * h->nDir * (h->nc + 1);
  CUDA_OK(h, cudaMemcpyAsync(out, h->d_intByComp, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return I3RC_SUCCESS;
}

void i3rc_get_counters(const i3rc_integrator* h, i3rc_counters* c) {
  if (h && c) *c = h->counters;
}

int i3rc_copy_Integrator(const i3rc_integrator* s, i3rc_integrator** out) {
  if (!s || !out) return I3RC_FAILURE;
  *out = nullptr;
  cudaSetDevice(s->device);
  // rebuild from the device-resident dense arrays (a full copy, unlike MCRT:1082-1253 which forgets the
  // intensity-limiting members: quirk Q6)
  const size_t ncell = (size_t)s->nx * s->ny * s->nz;
  cudaStreamSynchronize(s->stream);  // (a profile swap or a batch may still be running on the source's stream)
  i3rc_integrator* h = make_handle();
  h->nx = s->nx;
  h->ny = s->ny;
  h->nz = s->nz;
  h->nc = s->nc;
  h->xe = s->xe;
  h->ye = s->ye;
  h->ze = s->ze;
  h->maxPfIndex = s->maxPfIndex;
  h->absorbing = s->absorbing;
  h->splitLayers = s->splitLayers;
  auto fields = [&]() -> int {  // device to device, no trip through the host
    CUDA_OK(h, cudaMalloc(&h->d_ext, sizeof(float) * ncell));
    CUDA_OK(h, cudaMalloc(&h->d_cum, sizeof(float) * ncell * s->nc));
    CUDA_OK(h, cudaMalloc(&h->d_ssa, sizeof(float) * ncell * s->nc));
    CUDA_OK(h, cudaMalloc(&h->d_pf, sizeof(int) * ncell * s->nc));
    CUDA_OK(h, cudaMemcpyAsync(h->d_ext, s->d_ext, sizeof(float) * ncell, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_OK(h, cudaMemcpyAsync(h->d_cum, s->d_cum, sizeof(float) * ncell * s->nc, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_OK(h, cudaMemcpyAsync(h->d_ssa, s->d_ssa, sizeof(float) * ncell * s->nc, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_OK(h, cudaMemcpyAsync(h->d_pf, s->d_pf, sizeof(int) * ncell * s->nc, cudaMemcpyDeviceToDevice, h->stream));
    if (s->d_extRaw) {
      CUDA_OK(h, cudaMalloc(&h->d_extRaw, sizeof(float) * ncell * s->nc));
      CUDA_OK(h, cudaMemcpyAsync(h->d_extRaw, s->d_extRaw, sizeof(float) * ncell * s->nc, cudaMemcpyDeviceToDevice, h->stream));
    }
    return finish_new_integrator(h);
  };
  int rc = fields();
  if (rc == I3RC_FAILURE) {
    g_message = h->message;
    i3rc_finalize_Integrator(h);
    return rc;
  }
  for (int c = 0; c < s->nc; c++) {
    const HostTable& t = s->tables[c];
    if (t.kind == 0) continue;
    i3rc_phase_table pt{t.kind, t.nEntries, t.offsets.data(), t.coefs.data(), t.nAngles, t.angles.data(), t.values.data()};
    i3rc_set_phase_table(h, c, &pt);
  }
  i3rc_params p;
  memset(&p, 0, sizeof(p));
  std::vector<float> mus(s->nDir), phis(s->nDir);
  p.present = I3RC_P_minForwardTableSize | I3RC_P_minInverseTableSize | I3RC_P_useRayTracing | I3RC_P_useRussianRoulette |
              I3RC_P_useRussianRouletteForIntensity | I3RC_P_zetaMin | I3RC_P_useHybridPhaseFunsForIntenCalcs |
              I3RC_P_hybridPhaseFunWidth | I3RC_P_numOrdersOrigPhaseFunIntenCalcs | I3RC_P_limitIntensityContributions |
              I3RC_P_maxIntensityContribution;
  p.minForwardTableSize = s->minForwardTableSize;
  p.minInverseTableSize = s->minInverseTableSize;
  p.useRayTracing = s->useRayTracing;
  p.useRussianRoulette = s->useRussianRoulette;
  p.useRussianRouletteForIntensity = s->useRRIntensity;
  p.zetaMin = s->zetaMin;
  p.useHybridPhaseFunsForIntenCalcs = s->useHybrid;
  p.hybridPhaseFunWidth = s->hybridWidth;
  p.numOrdersOrigPhaseFunIntenCalcs = s->numOrdersOrig;
  p.limitIntensityContributions = s->limitContrib;
  p.maxIntensityContribution = s->maxContrib;
  if (s->useSurfaceBDRF) {
    p.present |= I3RC_P_surfaceBDRF;
    p.surf_nx = s->surf_nx;
    p.surf_ny = s->surf_ny;
    p.surf_x = s->surf_x.data();
    p.surf_y = s->surf_y.data();
    p.surf_params = s->surf_p.data();
  } else {
    p.present |= I3RC_P_surfaceAlbedo;
    p.surfaceAlbedo = s->surfaceAlbedo;
  }
  rc = i3rc_specifyParameters(h, &p);
  if (rc != I3RC_FAILURE && s->computeIntensity && s->nDir > 0) {
    // directions are stored as cosines; hand them over directly
    h->nDir = s->nDir;
    h->dirs = s->dirs;
    size_t ncol = (size_t)h->nx * h->ny;
    upload(&h->d_dirs, h->dirs.data(), h->dirs.size(), h->stream);
    h->leLBValid = false;
    cudaMalloc(&h->d_intensity, sizeof(float) * ncol * h->nDir);
    cudaMalloc(&h->d_intByComp, sizeof(float) * ncol * h->nDir * (h->nc + 1));
    cudaMemsetAsync(h->d_intensity, 0, sizeof(float) * ncol * h->nDir, h->stream);
    if (h->limitContrib) cudaMalloc(&h->d_excess, sizeof(float) * (size_t)(h->nc + 1) * h->nDir);
    cudaStreamSynchronize(h->stream);
    h->computeIntensity = true;
  }
  h->trackByComponent = s->trackByComponent;
  h->blockSize = s->blockSize;
  h->blocksPerSM = s->blocksPerSM;
  h->kSteps = s->kSteps;
  h->eventThreshold = s->eventThreshold;
  h->birthMin = s->birthMin;
  h->birthLow = s->birthLow;
  h->slots80 = s->slots80;
  h->residentBlocks = s->residentBlocks;
  h->poolShape = s->poolShape;
  h->minRunning = s->minRunning;
  h->splitLayers = s->splitLayers;
  h->message.clear();
  *out = h;
  return I3RC_SUCCESS;
}

// ---- probes --------------------------------------------------------------------------------------------
int i3rc_trace_rays(i3rc_integrator* h, int n, const float* pos, const float* dir, const float* tauLimit, float* tauOut,
                    float* posOut, int32_t* idxOut) {
  if (!h || !h->readyToCompute || n < 1) return I3RC_FAILURE;
  CUDA_OK(h, cudaSetDevice(h->device));
  float *d_pos = nullptr, *d_dir = nullptr, *d_lim = nullptr, *d_tau = nullptr, *d_po = nullptr;
  int* d_idx = nullptr;
  int rc = [&]() -> int {
    CUDA_OK(h, upload(&d_pos, pos, (size_t)3 * n, h->stream));
    CUDA_OK(h, upload(&d_dir, dir, (size_t)3 * n, h->stream));
    if (tauLimit) CUDA_OK(h, upload(&d_lim, tauLimit, (size_t)n, h->stream));
    CUDA_OK(h, cudaMalloc(&d_tau, sizeof(float) * n));
    CUDA_OK(h, cudaMalloc(&d_po, sizeof(float) * 3 * n));
    CUDA_OK(h, cudaMalloc(&d_idx, sizeof(int) * 3 * n));
    if (h->xyRegular && h->zRegular && h->nzc) {
      // the stepping code of the production kernel for regular grids with a layer table (uniform slabs crossed in one go)
      ProblemT<true, false, true> p;
      fill_problem(h, p);
      k_trace_rays<<<(n + 127) / 128, 128, 0, h->stream>>>(p, n, d_pos, d_dir, d_lim, d_tau, d_po, d_idx);
    } else {
      ProblemDyn p;
      fill_problem(h, p);
      k_trace_rays<<<(n + 127) / 128, 128, 0, h->stream>>>(p, n, d_pos, d_dir, d_lim, d_tau, d_po, d_idx);
    }
    h->otherLaunches++;
    CUDA_OK(h, cudaGetLastError());
    CUDA_OK(h, cudaMemcpyAsync(tauOut, d_tau, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    if (posOut) CUDA_OK(h, cudaMemcpyAsync(posOut, d_po, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    if (idxOut) CUDA_OK(h, cudaMemcpyAsync(idxOut, d_idx, sizeof(int) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return I3RC_SUCCESS;
  }();
  dfree(d_pos);
  dfree(d_dir);
  dfree(d_lim);
  dfree(d_tau);
  dfree(d_po);
  dfree(d_idx);
  return rc;
}

static int probe_1d(i3rc_integrator* h, const float* table, int nSteps, int n, const float* in, float* out, bool inverse) {
  float *d_in = nullptr, *d_out = nullptr;
  int rc = [&]() -> int {
    CUDA_OK(h, upload(&d_in, in, (size_t)n, h->stream));
    CUDA_OK(h, cudaMalloc(&d_out, sizeof(float) * n));
    if (inverse)
      k_sample_angles<<<(n + 127) / 128, 128, 0, h->stream>>>(table, nSteps, n, d_in, d_out);
    else
      k_lookup_phase<<<(n + 127) / 128, 128, 0, h->stream>>>(table, nSteps, n, d_in, d_out);
    h->otherLaunches++;
    CUDA_OK(h, cudaGetLastError());
    CUDA_OK(h, cudaMemcpyAsync(out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return I3RC_SUCCESS;
  }();
  dfree(d_in);
  dfree(d_out);
  return rc;
}
int i3rc_sample_scattering_angles(i3rc_integrator* h, int comp, int entry, int n, const float* xi, float* theta) {
  if (!h || comp < 0 || comp >= h->nc || !h->inv[comp].d || entry < 0 || entry >= h->inv[comp].nEntries) return I3RC_FAILURE;
  return probe_1d(h, h->inv[comp].d + (size_t)entry * h->inv[comp].nSteps, h->inv[comp].nSteps, n, xi, theta, true);
}
int i3rc_lookup_phase_function(i3rc_integrator* h, int comp, int entry, int which, int n, const float* angles, float* out) {
  if (!h || comp < 0 || comp >= h->nc) return I3RC_FAILURE;
  DevMatrix& m = which == 1 ? h->fwd[comp] : h->fwdOrig[comp];
  if (!m.d || entry < 0 || entry >= m.nEntries) return I3RC_FAILURE;
  return probe_1d(h, m.d + (size_t)entry * m.nSteps, m.nSteps, n, angles, out, false);
}

// ---- device probes of the random-number stream and of next_direct (no handle needed) -----------------------------
int i3rc_probe_philox(uint32_t key0, uint32_t key1, int n, const uint64_t* photon, const uint32_t* block, uint32_t* raw4,
                      float* u4) {
  if (!have_device() || n < 1 || !photon || !block || !raw4 || !u4) return I3RC_FAILURE;
  unsigned long long* d_ph = nullptr;
  uint32_t *d_bl = nullptr, *d_raw = nullptr;
  float* d_u = nullptr;
  bool ok = cudaMalloc(&d_ph, 8 * (size_t)n) == cudaSuccess && cudaMalloc(&d_bl, 4 * (size_t)n) == cudaSuccess &&
            cudaMalloc(&d_raw, 16 * (size_t)n) == cudaSuccess && cudaMalloc(&d_u, 16 * (size_t)n) == cudaSuccess;
  if (ok) {
    cudaMemcpy(d_ph, photon, 8 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_bl, block, 4 * (size_t)n, cudaMemcpyHostToDevice);
    k_probe_philox<<<(n + 127) / 128, 128>>>(key0, key1, n, d_ph, d_bl, d_raw, d_u);
    ok = cudaMemcpy(raw4, d_raw, 16 * (size_t)n, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(u4, d_u, 16 * (size_t)n, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  cudaFree(d_ph), cudaFree(d_bl), cudaFree(d_raw), cudaFree(d_u);
  return ok ? I3RC_SUCCESS : I3RC_FAILURE;
}
int i3rc_probe_next_direct(uint32_t key0, uint32_t key1, int n, const uint64_t* photon, const uint32_t* block,
                           const float* direction, const float* cosine, float* newDirection, uint32_t* blocksUsed) {
  if (!have_device() || n < 1 || !photon || !block || !direction || !cosine || !newDirection || !blocksUsed) return I3RC_FAILURE;
  unsigned long long* d_ph = nullptr;
  uint32_t *d_bl = nullptr, *d_used = nullptr;
  float *d_S = nullptr, *d_c = nullptr, *d_o = nullptr;
  bool ok = cudaMalloc(&d_ph, 8 * (size_t)n) == cudaSuccess && cudaMalloc(&d_bl, 4 * (size_t)n) == cudaSuccess &&
            cudaMalloc(&d_used, 4 * (size_t)n) == cudaSuccess && cudaMalloc(&d_S, 12 * (size_t)n) == cudaSuccess &&
            cudaMalloc(&d_c, 4 * (size_t)n) == cudaSuccess && cudaMalloc(&d_o, 12 * (size_t)n) == cudaSuccess;
  if (ok) {
    cudaMemcpy(d_ph, photon, 8 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_bl, block, 4 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_S, direction, 12 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_c, cosine, 4 * (size_t)n, cudaMemcpyHostToDevice);
    k_probe_next_direct<<<(n + 127) / 128, 128>>>(key0, key1, n, d_ph, d_bl, d_S, d_c, d_o, d_used);
    ok = cudaMemcpy(newDirection, d_o, 12 * (size_t)n, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(blocksUsed, d_used, 4 * (size_t)n, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  cudaFree(d_ph), cudaFree(d_bl), cudaFree(d_used), cudaFree(d_S), cudaFree(d_c), cudaFree(d_o);
  return ok ? I3RC_SUCCESS : I3RC_FAILURE;
}

// ---- roofline ceilings measured on this GPU (the bound of C1-C4 is the L2 random-gather rate) ----
// Random 4-byte gathers over `bytes` of device memory (rounded down to a power of two): *gathersPerSec = loads per second,
// best of `repeats` launches timed with CUDA events.
int i3rc_measure_gather_rate(size_t bytes, int repeats, double* gathersPerSec) {
  if (!have_device() || !gathersPerSec || bytes < 4096) return I3RC_FAILURE;
  size_t words = 1;
  while (words * 2 * sizeof(float) <= bytes) words *= 2;
  if (words > ((size_t)1 << 25)) words = (size_t)1 << 25;  // (the index is 25 bits of the generator's state)
  float *d = nullptr, *sink = nullptr;
  if (cudaMalloc(&d, words * sizeof(float)) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) {
    cudaFree(d);
    return I3RC_FAILURE;
  }
  cudaMemset(d, 0, words * sizeof(float));
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 256, unr = 8;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0.0;
  for (int r = 0; r < std::max(repeats, 1) + 1; r++) {  // (the first launch warms the cache up)
    cudaEventRecord(a);
    k_gather_bench<8><<<blocks, threads>>>(d, (uint32_t)(words - 1), iters, sink);
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) break;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    if (r > 0 && ms > 0.0f) best = std::max(best, (double)blocks * threads * iters * unr / (ms * 1e-3));
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  cudaFree(sink);
  *gathersPerSec = best;
  return best > 0.0 ? I3RC_SUCCESS : I3RC_FAILURE;
}
// Warp instructions per second the SMs can issue (independent FMAs from a full complement of warps)
int i3rc_measure_issue_rate(int repeats, double* warpInstPerSec) {
  if (!have_device() || !warpInstPerSec) return I3RC_FAILURE;
  float* sink = nullptr;
  if (cudaMalloc(&sink, 4) != cudaSuccess) return I3RC_FAILURE;
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0.0;
  for (int r = 0; r < std::max(repeats, 1) + 1; r++) {
    cudaEventRecord(a);
    k_issue_bench<<<blocks, threads>>>(iters, sink);
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) break;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    if (r > 0 && ms > 0.0f) best = std::max(best, (double)blocks * (threads / 32) * iters * 128.0 / (ms * 1e-3));
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(sink);
  *warpInstPerSec = best;
  return best > 0.0 ? I3RC_SUCCESS : I3RC_FAILURE;
}

// ---- batch moments -------------------------------------------------------------------------------------
int i3rc_stats_reset(i3rc_integrator* h, int with_volume) {
  if (!h || !h->readyToCompute) return I3RC_FAILURE;
  CUDA_OK(h, cudaSetDevice(h->device));
  int nD = h->computeIntensity ? h->nDir : 0;
  StatsLayout L = stats_layout(h, nD, with_volume != 0);
  if (h->statsN != L.total) {
    dfree(h->d_stats);
    CUDA_OK(h, cudaMalloc(&h->d_stats, sizeof(double) * L.total));
    h->statsN = L.total;
  }
  h->statsVolume = with_volume != 0;
  h->statsNDir = nD;
  CUDA_OK(h, cudaMemsetAsync(h->d_stats, 0, sizeof(double) * L.total, h->stream));
  return I3RC_SUCCESS;
}

int i3rc_stats_accumulate(i3rc_integrator* h) {
  if (!h || !h->d_stats) return h ? fail(h, "stats_accumulate: call stats_reset first") : I3RC_FAILURE;
  if (h->statsNDir != (h->computeIntensity ? h->nDir : 0))  // (the moment buffer was laid out for other directions)
    return fail(h, "stats_accumulate: the intensity directions changed since stats_reset; call stats_reset again");
  int nD = h->statsNDir;
  StatsLayout L = stats_layout(h, nD, h->statsVolume);
  size_t ncol = (size_t)h->nx * h->ny, ncell = ncol * h->nz;
  int rows = 3 + h->nz + nD;
  if (ensure_scratch(h, rows) != I3RC_SUCCESS) return I3RC_FAILURE;
  double* S = h->d_stats;
  cudaStream_t st = h->stream;
  k_slab_sums<<<1, 256, 0, st>>>(h->d_fluxUp, ncol, h->d_scratch + 0);
  k_slab_sums<<<1, 256, 0, st>>>(h->d_fluxDown, ncol, h->d_scratch + 1);
  k_slab_sums<<<1, 256, 0, st>>>(h->d_fluxAbs, ncol, h->d_scratch + 2);
  k_slab_sums<<<h->nz, 256, 0, st>>>(h->d_volAbs, ncol, h->d_scratch + 3);
  if (nD) k_slab_sums<<<nD, 256, 0, st>>>(h->d_intensity, ncol, h->d_scratch + 3 + h->nz);
  double inv = 1.0 / (double)ncol;
  for (int i = 0; i < 3; i++) k_moments_mean<<<1, 32, 0, st>>>(h->d_scratch + i, 1, inv, S + L.mean[i], S + L.mean[i] + 1);
  k_moments_mean<<<(h->nz + 127) / 128, 128, 0, st>>>(h->d_scratch + 3, h->nz, inv, S + L.prof, S + L.prof + h->nz);
  if (nD) k_moments_mean<<<1, 32, 0, st>>>(h->d_scratch + 3 + h->nz, nD, inv, S + L.mrad, S + L.mrad + nD);
  unsigned gcol = (unsigned)((ncol + 255) / 256);
  k_moments_f<<<gcol, 256, 0, st>>>(h->d_fluxUp, ncol, S + L.flux[0], S + L.flux[0] + ncol);
  k_moments_f<<<gcol, 256, 0, st>>>(h->d_fluxDown, ncol, S + L.flux[1], S + L.flux[1] + ncol);
  k_moments_f<<<gcol, 256, 0, st>>>(h->d_fluxAbs, ncol, S + L.flux[2], S + L.flux[2] + ncol);
  if (nD) k_moments_f<<<(unsigned)((ncol * nD + 255) / 256), 256, 0, st>>>(h->d_intensity, ncol * nD, S + L.rad, S + L.rad + ncol * nD);
  if (h->statsVolume) k_moments_f<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(h->d_volAbs, ncell, S + L.vol, S + L.vol + ncell);
  h->otherLaunches += 11 + (nD ? 3 : 0) + (h->statsVolume ? 1 : 0);
  CUDA_OK(h, cudaGetLastError());
  return I3RC_SUCCESS;
}

int i3rc_stats_device_buffer(i3rc_integrator* h, void** dev_ptr, int64_t* n_doubles) {
  if (!h || !h->d_stats) return I3RC_FAILURE;
  if (dev_ptr) *dev_ptr = h->d_stats;
  if (n_doubles) *n_doubles = (int64_t)h->statsN;
  return I3RC_SUCCESS;
}

int i3rc_stats_report(i3rc_integrator* h, double solarFlux, int numBatches, const i3rc_stats_out* out) {
  if (!h || !h->d_stats || !out) return I3RC_FAILURE;
  if (numBatches < 2) return fail(h, "stats_report: at least two batches are needed (monteCarloDriver.f95:264)");
  CUDA_OK(h, cudaSetDevice(h->device));
  int nD = h->statsNDir;
  StatsLayout L = stats_layout(h, nD, h->statsVolume);
  size_t ncol = (size_t)h->nx * h->ny, ncell = ncol * h->nz;
  double* d_tmp = nullptr;
  size_t maxBlock = std::max({ncol * (size_t)std::max(nD, 1), (size_t)h->nz, h->statsVolume ? ncell : (size_t)1, (size_t)8});
  CUDA_OK(h, cudaMalloc(&d_tmp, sizeof(double) * 2 * maxBlock));
  auto emit = [&](size_t off, size_t n, double* dst) -> int {
    if (!dst || n == 0) return I3RC_SUCCESS;
    k_stats_finish<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_stats + off, h->d_stats + off + n, n, solarFlux,
                                                                      numBatches, d_tmp, d_tmp + n);
    CUDA_OK(h, cudaMemcpyAsync(dst, d_tmp, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    h->otherLaunches++;
    return I3RC_SUCCESS;
  };
  int rc = I3RC_SUCCESS;
  double* means[3] = {out->meanFluxUp, out->meanFluxDown, out->meanFluxAbsorbed};
  double* fluxes[3] = {out->fluxUp, out->fluxDown, out->fluxAbsorbed};
  for (int i = 0; i < 3 && rc == I3RC_SUCCESS; i++) rc = emit(L.mean[i], 1, means[i]);
  for (int i = 0; i < 3 && rc == I3RC_SUCCESS; i++) rc = emit(L.flux[i], ncol, fluxes[i]);
  if (rc == I3RC_SUCCESS) rc = emit(L.prof, h->nz, out->absorbedProfile);
  if (rc == I3RC_SUCCESS && nD) rc = emit(L.rad, ncol * nD, out->radiance);
  if (rc == I3RC_SUCCESS && nD) rc = emit(L.mrad, nD, out->meanRadiance);
  if (rc == I3RC_SUCCESS && h->statsVolume) rc = emit(L.vol, ncell, out->absorbedVolume);
  cudaFree(d_tmp);
  return rc;
}

int i3rc_run_batches(i3rc_integrator* h, const i3rc_photon_source* src, int32_t iseed, int seedOrder, int batchBegin,
                     int nBatches) {
  SourceDev sd;
  int rc = prepare_compute(h, src, sd);
  if (rc != I3RC_SUCCESS) return rc;
  if (!h->d_stats || h->statsNDir != (h->computeIntensity ? h->nDir : 0)) {  // no moments yet, or laid out for other directions
    rc = i3rc_stats_reset(h, h->d_stats ? (int)h->statsVolume : 0);
    if (rc != I3RC_SUCCESS) return rc;
  }
  CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * CNT_N, h->stream));
  for (int b = 0; b < nBatches; b++) {
    int32_t seed[2];
    seed[0] = seedOrder == 0 ? iseed : batchBegin + b;
    seed[1] = seedOrder == 0 ? batchBegin + b : iseed;
    uint32_t k0, k1;
    seed_to_key(seed, 2, k0, k1);
    rc = run_one_batch(h, sd, k0, k1);
    if (rc != I3RC_SUCCESS) return rc;
    rc = i3rc_stats_accumulate(h);
    if (rc != I3RC_SUCCESS) return rc;
  }
  return fetch_counters(h);
}

// ---- NCCL (dlopen'ed so that the library neither links NCCL nor clashes with the copy torch bundles) -------
namespace {
struct NcclId {
  char internal[128];
};
typedef int (*fn_getid)(NcclId*);
typedef int (*fn_init)(void**, int, NcclId, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
void* nccl_lib() {
  static void* lib = nullptr;
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  return lib;
}
}  // namespace

int i3rc_comm_unique_id(char* id128) {
  void* lib = nccl_lib();
  if (!lib || !id128) {
    g_message = "comm_unique_id: libnccl.so.2 not found";
    return I3RC_FAILURE;
  }
  fn_getid f = (fn_getid)dlsym(lib, "ncclGetUniqueId");
  NcclId id;
  if (!f || f(&id) != 0) return I3RC_FAILURE;
  memcpy(id128, id.internal, 128);
  return I3RC_SUCCESS;
}
int i3rc_comm_init(i3rc_integrator* h, int nRanks, int rank, const char* id128) {
  if (!h || !id128) return I3RC_FAILURE;
  void* lib = nccl_lib();
  if (!lib) return fail(h, "comm_init: libnccl.so.2 not found");
  fn_init f = (fn_init)dlsym(lib, "ncclCommInitRank");
  if (!f) return fail(h, "comm_init: ncclCommInitRank not found");
  CUDA_OK(h, cudaSetDevice(h->device));
  NcclId id;
  memcpy(id.internal, id128, 128);
  void* comm = nullptr;
  if (f(&comm, nRanks, id, rank) != 0) return fail(h, "comm_init: ncclCommInitRank failed");
  h->nccl = comm;
  h->ncclLib = lib;
  return I3RC_SUCCESS;
}
int i3rc_stats_allreduce(i3rc_integrator* h) {
  if (!h || !h->d_stats) return I3RC_FAILURE;
  if (!h->nccl) return I3RC_SUCCESS;  // single process: identity, like multipleProcesses_nompi.f95
  fn_allreduce f = (fn_allreduce)dlsym(h->ncclLib, "ncclAllReduce");
  if (!f) return fail(h, "stats_allreduce: ncclAllReduce not found");
  // ncclFloat64 = 8, ncclSum = 0
  if (f(h->d_stats, h->d_stats, h->statsN, 8, 0, h->nccl, h->stream) != 0) return fail(h, "stats_allreduce: ncclAllReduce failed");
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return I3RC_SUCCESS;
}
int i3rc_comm_finalize(i3rc_integrator* h) {
  if (!h || !h->nccl) return I3RC_SUCCESS;
  fn_destroy f = (fn_destroy)dlsym(h->ncclLib, "ncclCommDestroy");
  if (f) f(h->nccl);
  h->nccl = nullptr;
  return I3RC_SUCCESS;
}

// ---- streams, timing, tuning ---------------------------------------------------------------------------
int i3rc_synchronize(i3rc_integrator* h) {
  if (!h) return I3RC_FAILURE;
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  return I3RC_SUCCESS;
}
void* i3rc_stream(i3rc_integrator* h) { return h ? (void*)h->stream : nullptr; }

int i3rc_get_timing(i3rc_integrator* h, double* trace_ms, int64_t* trace_launches, int64_t* other_launches) {
  if (!h) return I3RC_FAILURE;
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i + 1 < h->evUsed; i += 2) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess) h->traceMs += ms;
  }
  h->evUsed = 0;
  if (trace_ms) *trace_ms = h->traceMs;
  if (trace_launches) *trace_launches = h->traceLaunches;
  if (other_launches) *other_launches = h->otherLaunches;
  return I3RC_SUCCESS;
}
int i3rc_reset_timing(i3rc_integrator* h) {
  if (!h) return I3RC_FAILURE;
  CUDA_OK(h, cudaStreamSynchronize(h->stream));
  h->evUsed = 0;
  h->traceMs = 0.0;
  h->traceLaunches = h->otherLaunches = 0;
  return I3RC_SUCCESS;
}
// what new_Integrator / the next launch decided: 0 = number of layers stored in 3-D when the uniform layers are kept out of
// the extinction field (0: all layers stored), 1 = floats of tallies staged per warp in shared memory (0: global atomics)
int i3rc_get_layout(i3rc_integrator* h, int what) {
  if (!h || !h->readyToCompute) return -1;
  if (what == 0) return h->nzc;
  if (what == 1) {
    Problem p;
    fill_problem(h, p);
    return p.tsmN;
  }
  if (what == 2) return h->d_extJ ? (int)(h->codedFraction * 1000.0 + 0.5) : 0;  // per mille of cells with an empty-space code
  return -1;
}

int i3rc_set_tuning(i3rc_integrator* h, const char* key, int value) {
  if (!h || !key) return I3RC_FAILURE;
  std::string k(key);
  if (k == "block_size" && value == 128)
    h->blockSize = value;
  else if (k == "blocks_per_sm" && value >= 0)
    h->blocksPerSM = value;
  else if (k == "steps_per_event_phase" && value >= 1)
    h->kSteps = value;
  else if (k == "resident_blocks" && (value == 0 || (value >= 4 && value <= 8)))
    h->residentBlocks = value;
  else if (k == "split_layers")
    h->splitLayers = value;  // takes effect at the next new_Integrator / copy_Integrator
  else if (k == "pad_smem" && value >= 0 && value <= 16384)
    h->padSmem = value;  // experiment: unused dynamic shared memory (shrinks the L1 share of the SM)
  else if (k == "min_running" && value >= 0 && value <= 32)
    h->minRunning = value;
  else if (k == "pool_shape" && value >= 0 && value <= 6)
    h->poolShape = value;
  else if (k == "event_threshold" && value >= 1 && value <= 32)
    h->eventThreshold = value;
  else if (k == "birth_min" && value >= 1 && value <= 32)
    h->birthMin = value;
  else if (k == "birth_low" && value >= 0 && value <= 32)
    h->birthLow = value;
  else if (k == "slots_80" && (value == 0 || value == 1))
    h->slots80 = value;
  else if (k == "track_by_component")
    h->trackByComponent = value != 0;
  else if (k == "slab_jump" && (value == 0 || value == 1))
    h->slabJump = value;  // 0: runs of horizontally uniform layers are walked cell by cell
  else if (k == "l2_persist" && (value == 0 || value == 1)) {
    h->l2Persist = value;
    h->l2WindowFor = nullptr;
  } else if (k == "tables_in_smem" && (value == 0 || value == 1))
    h->tablesInSmem = value;
  else if (k == "debug_zero_strides" && (value == 0 || value == 1))
    h->debugZeroStrides = value;
  else if (k == "le_early_exit" && (value == 0 || value == 1))
    h->leEarlyExit = value;
  else if (k == "le_upper_bound" && (value == 0 || value == 1))
    h->leUpperBound = value;  // 0: rays that are certain to survive the roulette are traced all the same
  else if (k == "le_lower_bound" && value >= 0 && value <= 2)
    h->leLowerBound = value;  // 0: local-estimate rays are traced even when a lower bound says they cannot contribute
  else if (k == "vertical_shortcut" && (value == 0 || value == 1))
    h->verticalShortcut = value;  // 0: straight-up local-estimate rays are traced like the others
  else if (k == "skip_empty" && (value == 0 || value == 1)) {
    h->skipEmpty = value;  // 0: drop the coded copy of the field (takes effect at once); 1: at the next new_Integrator
    if (!value) dfree(h->d_extJ);
  } else if (k == "stage_max_columns" && value >= 0)
    h->stageMaxColumns = value;
  else if (k == "replica_columns" && value >= 0)
    h->replicaColumns = value;  // 0: no copies of the tallies in global memory
  else if (k == "stage_tallies" && value >= 0 && value <= 4096)
    h->stageTallies = value;  // floats of shared memory per warp for staged tallies; 0 = global atomics only
  else
    return fail(h, "set_tuning: unknown key or bad value");
  return I3RC_SUCCESS;
}

}  // extern "C"
