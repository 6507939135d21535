// __global__ kernels of the integrator (sm_100a).  The photon transport kernel is the hot path; the rest are
// the small setup / normalisation / reporting kernels around it.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "transport.cuh"

namespace i3rc {

#ifndef I3RC_STEP_UNROLL
#define I3RC_STEP_UNROLL 1  // pairs of cell crossings per unrolled loop body of a trace round
#endif

// ---- K1: persistent, warp-cooperative photon transport ---------------------------------------------------------
// grid = (#SMs x resident blocks).  Lanes are NOT tied to photons.  Every warp owns, in shared memory,
//   * a pool of NSLOT photon slots (position, direction, weight, Philox counter, raw end of the last segment: 56 bytes),
//   * a ring of QCAP ray TASKS (40 bytes each): a photon's next path segment, or one local-estimate ray,
//   * the list of slots whose path segment has ended and whose event (boundary / collision) is due,
// and alternates, as a warp-uniform state machine, between
//   TRACE rounds:  up to STEPS cell crossings for every lane that holds a ray (the round ends early when fewer than
//                  minRunning lanes still run); then, for all lanes at once, finished rays are closed (a segment's raw
//                  ray goes back to its slot, which joins the event list; a local-estimate ray is tallied) and idle
//                  lanes pop the next tasks from the ring;
//   EVENT batches: when 32 events are due (or the ring has run dry and at most lowWater lanes trace) lane i takes the
//                  i-th due slot: the raw ray becomes the event point, boundary or collision handling, then a warp-uniform loop over the radiance directions in which every lane
//                  turns its local-estimate ray into a task, then roulette + scattering, and the task of the next path
//                  segment.  A slot whose photon has finished waits among the warp's EMPTY slots; when birthMin of them
//                  are empty they join the due events and one batch starts all their photons (one warp-aggregated
//                  atomicAdd on the device photon counter, the birth code run by many lanes).
// So both the cell-crossing loop and the event code run with (nearly) full warps, whatever the individual photons do.
// A batch that finds the ring full is suspended between two directions and resumed after more trace rounds.
// Variants (template flags): TSM = the tallies of a few-column domain are staged per warp in shared memory and committed
// warp-aggregated (warp_tally, flush_staged_tallies); JUMP = rays use the empty-space codes of the gather field
// (transport.cuh, ray_advance_far; measured slower, off by default); TABSM = one 16-warp block per SM with the
// phase-function tables staged in shared memory (measured slower, an experiment); NSLOT = 80 (domains of many columns and
// few directions) keeps the state of a suspended batch in global memory instead of shared memory; SPLIT (layer table) runs
// ONE copy of the step per loop iteration instead of a pair.  A radiance direction that points
// straight up is never traced: make_le_task works its contribution out from the column's suffix sums.
// The physics functions are the ones of transport.cuh; the per-photon Philox streams make the result independent of
// which lane traces which ray (up to float summation order in the tallies).
template <int NSLOT>
struct SlotPool {  // structure of arrays: lane i touches slot[k] of every array
  uint32_t xy[NSLOT];   // cx | cy << 16
  uint32_t zs[NSLOT];   // cz | segDone << 16 (| SLOT_RAW: xy / zs / f* hold the raw ray, see below)
  float fx[NSLOT], fy[NSLOT], fz[NSLOT];
  float ux[NSLOT + MAX_DIRS], uy[NSLOT + MAX_DIRS], uz[NSLOT + MAX_DIRS];  // [NSLOT + d]: radiance direction d
  float w[NSLOT];
  int order[NSLOT];
  uint32_t id[NSLOT];     // photon number inside this launch
  uint32_t block[NSLOT];  // Philox blocks consumed so far
  // A path segment that has ended leaves its ray here as it is (zs bit 24): fx/fy/fz = path lengths left to the next
  // faces (signed, see Lane), xy/zs = the per-axis cell counters + how it ended, sp/spd = length of the pending
  // cell and the distance into it at which the segment ends.  The event batch turns that into the event point, with
  // all its lanes, instead of the one or two lanes that close a segment in any given trace round.
  float sp[NSLOT], spd[NSLOT];
};
constexpr uint32_t SLOT_RAW = 1u << 24;
#ifndef I3RC_SPLIT_SINGLE_STEP
#define I3RC_SPLIT_SINGLE_STEP 1  // 0: the layer-table kernels step in pairs like the others (A/B)
#endif

// everything a warp keeps in shared memory (8 KB with NSLOT = QCAP = 64, and with NSLOT = 80)
// (more than 64 slots: the scratch of suspended batches moves to global memory, Problem::susp, and its 896 bytes are the
//  sixteen extra slots -- the warp's block stays at 8 KB and the SM's L1 share unchanged)
template <int NSLOT>
struct SuspScratch {
  uint32_t w[NSLOT > 64 ? 1 : 7 * 32];  // per-lane state of a suspended event batch
};
template <int NSLOT, int QCAP>
struct WarpShared {
  LeTask task[QCAP];
  SlotPool<NSLOT> pool;
  SuspScratch<NSLOT> susp;
  uint32_t ray[4][32];   // per lane: what its ray is for (photon slot, or local-estimate parameters)
  uint32_t cnt[CNT_N];
  uint8_t pend[NSLOT];
};

// One tally increment per lane, committed by the whole warp (call it converged).  Domains of a few columns (planeParallel:
// one, the step cloud: 32) would otherwise send every increment of the GPU to a handful of addresses.  There, each
// warp owns a private copy of the tallies in shared memory (wt): increments of a warp that go to the same element are
// first summed over the lanes (match.any; a butterfly when it is the whole warp), one lane adds the sum without any
// atomic, and at the end of the kernel the block sums its warps' copies and issues one global atomic per element
// (flush_staged_tallies).  Larger domains (TSM = false, or a tally that is not staged) use global atomics directly.
template <bool TSM, class P>
__device__ __forceinline__ void warp_tally(const P& p, float* wt, bool has, int which, uint32_t off, float v) {
  if (TSM) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int s = -1;
    if (has) {
      const int o = p.tsmOff[which];
      if (o >= 0)
        s = o + (int)off;
      else
        atomicAdd(tally_ptr(p, which) + off, v);
    }
    const unsigned m = __ballot_sync(full, s >= 0);
    if (!m) return;
    const unsigned peers = __match_any_sync(full, s >= 0 ? s : -1 - lane);
    if (__all_sync(full, s < 0 || peers == m)) {  // one element for all: butterfly
      float x = s >= 0 ? v : 0.0f;
#pragma unroll
      for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(full, x, o);
      if (lane == __ffs(m) - 1) wt[s] += x;
    } else {  // groups of lanes with the same element: the first lane of a group collects
      float sum = 0.0f;
      unsigned rem = s >= 0 ? peers : 0u;
      while (__any_sync(full, rem != 0u)) {
        const int src = rem ? __ffs(rem) - 1 : lane;
        const float x = __shfl_sync(full, v, src);
        if (rem) sum += x;
        rem &= rem - 1u;
      }
      if (s >= 0 && lane == __ffs(peers) - 1) wt[s] += sum;
    }
    __syncwarp();
  } else if (has) {
    atomicAdd(tally_ptr(p, which) + off, v);
  }
}
// end of the kernel: block-level sum of the warps' private tallies, one global atomic per non-zero element
template <int BLOCK, class P>
__device__ __forceinline__ void flush_staged_tallies(const P& p, const float* sdyn) {
  __syncthreads();
  for (int i = threadIdx.x; i < p.tsmN; i += BLOCK) {
    float x = 0.0f;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++) x += sdyn[w * p.tsmN + i];
    if (x != 0.0f) {
      int which = 0, best = -1;
#pragma unroll
      for (int t = 0; t < 5; t++)
        if (p.tsmOff[t] >= 0 && p.tsmOff[t] <= i && p.tsmOff[t] > best) best = p.tsmOff[t], which = t;
      atomicAdd(tally_ptr(p, which) + (i - best), x);
    }
  }
}

// the 7 x 32 words of this warp in the scratch of suspended event batches (global memory; lane-private entries)
constexpr int SUSP_WORDS = 7 * 32;
template <int NSLOT, class P, class WS>
__device__ __forceinline__ uint32_t* susp_of(const P& p, WS& W, int warp) {
  if constexpr (NSLOT > 64)
    return p.susp + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * SUSP_WORDS;
  else
    return W.susp.w;
}

template <int BLOCK, bool REG, bool FAST, bool SPLIT, int MINB, int STEPS, int NSLOT, int QCAP, bool TSM = false, bool JUMP = false,
          bool TABSM = false>
__global__ void __launch_bounds__(BLOCK, MINB) k_transport(const ProblemT<REG, FAST, SPLIT, JUMP, TABSM> p, const int lowWater, const int minRunning,
                                                               const int births) {
  constexpr int NW = BLOCK / 32;
  extern __shared__ float s_dyn[];  // TSM: NW private copies of the staged tallies (Problem::tsmN floats each)
  using Tally = typename std::conditional<TSM, TallyLater, TallyNow>::type;
  constexpr int UNROLL = I3RC_STEP_UNROLL;
  static_assert(STEPS % (2 * UNROLL) == 0, "a round is a whole number of unrolled bodies");
  static_assert((QCAP & (QCAP - 1)) == 0 && QCAP >= 64, "ring size: power of two, room for one push of 32");
  static_assert(NSLOT >= 32 && NSLOT <= 255, "slot ids are bytes");
  // The warps' state is a static array for the usual blocks of 4 warps; the one-block-per-SM variant that also stages the
  // phase-function tables (TABSM: 16 warps + two 40 KB tables) carves everything out of dynamic shared memory.
  static_assert(!(TABSM && TSM), "the staging of tables and of tallies are not combined");
  __shared__ WarpShared<NSLOT, QCAP> s_warps[TABSM ? 1 : NW];
  WarpShared<NSLOT, QCAP>* const warps = TABSM ? reinterpret_cast<WarpShared<NSLOT, QCAP>*>(s_dyn) : s_warps;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  WarpShared<NSLOT, QCAP>& W = warps[warp];
  LeTask* q = W.task;
  SlotPool<NSLOT>& pool = W.pool;
  uint8_t* pend = W.pend;
  float* wt = s_dyn + (TSM ? warp * p.tsmN : 0);
  if constexpr (TABSM) {
    // one component, one table entry: the inverse and the forward table (40 KB each at 10001 steps) behind the warps' state
    float* tab = s_dyn + (sizeof(WarpShared<NSLOT, QCAP>) * NW + 15) / 16 * 4;
    const TableDesc T = p.tables[0];
    for (int i = threadIdx.x; i < T.nInv; i += BLOCK) tab[i] = __ldg(T.inv + i);
    for (int i = threadIdx.x; i < T.nFwd; i += BLOCK) tab[T.nInv + i] = __ldg(T.fwd + i);
    if (threadIdx.x == 0) {
      g_smTables = T;
      g_smTables.inv = tab;
      g_smTables.fwd = tab + T.nInv;
      g_smTables.fwdOrig = tab + T.nInv;
    }
    __syncthreads();
  }
  if (TSM) {
    for (int i = threadIdx.x; i < NW * p.tsmN; i += BLOCK) s_dyn[i] = 0.0f;
    __syncthreads();
  }

  // warp-uniform bookkeeping (registers): ring positions [head, tail), number of due events
  int head = 0, tail = 0, npend = NSLOT;
  const int birthMin = births & 255, birthLow = births >> 8;
  int nempty = 0;  // slots whose photon is finished and that wait for the next round of births (pend[NSLOT - 1 - i])
  bool exhausted = false;
  for (int k = lane; k < NSLOT; k += 32) {  // every slot starts by asking for a photon
    pend[k] = (uint8_t)k;
    pool.zs[k] = (uint32_t)DONE_NEW << 16;
  }
  if (lane < CNT_N) W.cnt[lane] = 0;
  if (lane < p.nDir) {  // (at most MAX_DIRS = 32 directions)
    pool.ux[NSLOT + lane] = __ldg(p.dirs + lane * DIR_STRIDE + 0);
    pool.uy[NSLOT + lane] = __ldg(p.dirs + lane * DIR_STRIDE + 1);
    pool.uz[NSLOT + lane] = __ldg(p.dirs + lane * DIR_STRIDE + 2);
  }
  __syncwarp();

  Lane R;  // the ray this lane is tracing
  R.cnt = W.cnt;
  R.done = DONE_IDLE;
  R.mode = MODE_PHOTON;
  R.nsteps = 0;
  R.slot = 0;

  // event batch control (kept across trace rounds); everything else of a suspended batch waits in shared memory
  int stage = 0;  // 0: no batch in progress, 1: local-estimate directions (next: dcur), 2: scattering + next segment
  int dcur = 0;

  for (;;) {
    // ================= EVENT batch =================
    // A batch starts when 32 events are due, or when the ring has run dry and few lanes are tracing.  It normally
    // runs to its end here; if the ring gets full it is suspended (state to shared memory) and resumed after the
    // next trace round.
    bool run = stage != 0;
    int eslot = 0;
    bool has = false, alive = false;
    if (stage == 0) {
      const int busy = __popc(__ballot_sync(full, R.done != DONE_IDLE));
      const int queued = tail - head;
      // Births in groups: enough empty slots (or nothing else to do) -> they go on top of the due events, so that the
      // next batch starts their photons with many lanes at once (one request to the device photon counter for all).
      if (nempty >= birthMin || (nempty > 0 && queued == 0 && busy <= birthLow)) {
        const int m = min(nempty, 32);
        const uint8_t s = lane < m ? pend[NSLOT - 1 - (nempty - m + lane)] : (uint8_t)0;
        __syncwarp();
        if (lane < m) pend[npend + lane] = s;
        __syncwarp();
        npend += m;
        nempty -= m;
      }
      if (npend >= 32 || (npend > 0 && queued == 0 && busy <= lowWater)) {
        const int k = min(32, npend);
        has = lane < k;
        eslot = has ? pend[npend - k + lane] : 0;
        npend -= k;
        run = true;
      } else if (npend == 0 && queued == 0 && busy == 0) {
        break;  // nothing in flight, nothing queued, nothing due: all slots are empty
      }
    }
    if (run) {
      if (stage != 0) {  // resuming a suspended batch
        const uint32_t w = susp_of<NSLOT>(p, W, warp)[6 * 32 + lane];
        eslot = (int)(w & 0xffu);
        has = (w & 0x100u) != 0;
      }
      Lane E;  // the photon whose event this lane is processing
      E.cnt = W.cnt;
      E.active = 0;
      E.comp = 0;
      E.pfi = 0;
      E.eCell = 0.0f;
      float a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;  // deviates of the event's block still to be used
      float c2 = 0.0f, c3 = 0.0f;             // second half of the block shared by two local-estimate directions
      if (has) {
        const uint32_t zs = pool.zs[eslot];
        E.segDone = (int)((zs >> 16) & 15u);
        if (E.segDone != DONE_NEW) {
          const uint32_t xy = pool.xy[eslot];
          E.fx = pool.fx[eslot];
          E.fy = pool.fy[eslot];
          E.fz = pool.fz[eslot];
          E.ux = pool.ux[eslot];
          E.uy = pool.uy[eslot];
          E.uz = pool.uz[eslot];
          if (zs & SLOT_RAW) {  // the segment's ray as it stopped -> the event point (MCRT:1721-1731, 1793-1804)
            Lane T;
            T.stx = E.ux >= 0.0f ? 1 : -1;  // (only the signs matter here)
            T.sty = E.uy >= 0.0f ? 1 : -1;
            T.stz = E.uz >= 0.0f ? 1 : -1;
            T.nax = -inv_abs(E.ux) * (REG ? p.dx : 1.0f);
            T.nay = -inv_abs(E.uy) * (REG ? p.dy : 1.0f);
            T.naz = -inv_abs(E.uz) * (REG ? p.dz : 1.0f);
            T.rx = E.fx;
            T.ry = E.fy;
            T.rz = E.fz;
            T.cntx = (int)(xy & 0xffffu);
            T.cnty = (int)(xy >> 16);
            T.cntz = (int)(zs & 0xffffu);
            T.sp = pool.sp[eslot];
            if (E.segDone == DONE_INSIDE) ray_undo_to(p, T, pool.spd[eslot]);
            ray_local(p, T, &E.fx, &E.fy, &E.fz);
            E.cx = ray_ix(p, T);
            E.cy = ray_iy(p, T);
            E.cz = ray_iz(p, T);
          } else {
            E.cx = (int)(xy & 0xffffu);
            E.cy = (int)(xy >> 16);
            E.cz = (int)(int16_t)(zs & 0xffffu);
          }
          E.w = pool.w[eslot];
          E.order = pool.order[eslot];
          E.rng.init((uint64_t)(p.firstPhoton + (long long)pool.id[eslot]));
          E.rng.block = pool.block[eslot];
          E.active = 1;
        }
      }
      if (stage == 0) {  // boundary / collision handling (MCRT:499-561, 581-649)
        I3RC_STAT(W, ST_BATCH, 1);
        I3RC_STAT(W, ST_BATCH_HAS, __popc(__ballot_sync(full, has)));
        alive = false;
        Tally tal;
        if constexpr (TSM) tal.n = 0;
        if (has && E.segDone != DONE_NEW) {
          float a0;
          E.rng.next4(p.key0, p.key1, a0, a1, a2, a3);
          alive = photon_event(p, E, a0, a1, tal) != 0;
        }
        if constexpr (TSM) {  // flux through the top / the surface, or absorption (column and cell): committed with the whole warp
          const TallyLater& t = tal;
          warp_tally<TSM>(p, wt, t.n > 0, t.w0, t.o0, t.v0);
          if (__any_sync(full, t.n > 1)) warp_tally<TSM>(p, wt, t.n > 1, t.w1, t.o1, t.v1);
        }
        dcur = 0;
        I3RC_STAT(W, ST_BATCH_ALIVE, __popc(__ballot_sync(full, alive)));
        stage = (p.computeIntensity && __any_sync(full, alive)) ? 1 : 2;
      } else {  // resume: the rest of the batch state comes back from shared memory
        const uint32_t* sv = susp_of<NSLOT>(p, W, warp) + lane;
        alive = (sv[6 * 32] & 0x200u) != 0;
        a1 = __uint_as_float(sv[0 * 32]);
        a2 = __uint_as_float(sv[1 * 32]);
        a3 = __uint_as_float(sv[2 * 32]);
        c2 = __uint_as_float(sv[3 * 32]);
        c3 = __uint_as_float(sv[4 * 32]);
        E.comp = (int)(sv[5 * 32] >> 16);
        E.pfi = (int)(sv[5 * 32] & 0xffffu);
        if (has && E.active) E.eCell = ext_at(p, E.cx, E.cy, E.cz);
      }
      if (stage == 1) {  // local-estimate tasks, one direction at a time (MCRT:1473-1569)
        while (dcur < p.nDir && QCAP - (tail - head) >= 32) {
          float c0 = c2, c1 = c3;
          if ((dcur & 1) == 0 && alive) E.rng.next4(p.key0, p.key1, c0, c1, c2, c3);  // one block per two directions
          LeTask t;
          const int what = alive ? make_le_task(p, E, dcur, c0, c1, t) : 0;
          const bool push = what == 1;
          const unsigned m = __ballot_sync(full, push);
          if (push) q[(tail + __popc(m & lt)) & (QCAP - 1)] = t;
          tail += __popc(m);
          I3RC_STAT(W, ST_LE_PUSH, __popc(m));
          if (((p.vertMask >> dcur) & 1u) || p.leUB) {  // contributions worked out on the spot (make_le_task returned 2)
            Tally tv;
            if constexpr (TSM) tv.n = 0;
            if (what == 2) tally_intensity_at(p, E, dcur, E.comp, (int)(t.xy >> 16) * p.nx + (int)(t.xy & 0xffffu), t.cw, tv);
            if constexpr (TSM) warp_tally<TSM>(p, wt, tv.n > 0, tv.w0, tv.o0, tv.v0);
          }
          dcur++;
        }
        if (dcur >= p.nDir) stage = 2;
      }
      if (stage == 2 && QCAP - (tail - head) >= 32) {
        bool go = false;
        float xiTau = a3;
        if (alive) {
          scatter_photon(p, E, a1, a2);  // roulette, new direction (MCRT:670-688)
          go = E.active != 0;
        }
        // Slots that come without a photon take the next photons of the batch.  A slot whose photon has just finished
        // waits among the empty ones until there are birthMin of them (above): the start of a photon's life -- a request
        // to the device counter, which the whole warp waits for, and a good 250 instructions -- is then run by many
        // lanes at once instead of by the one or two photons that finish in any given batch.
        const bool fresh = has && !go && !exhausted;
        const bool need = fresh && (E.segDone == DONE_NEW || birthMin <= 1);
        const unsigned md = __ballot_sync(full, fresh && !need);
        if (md) {
          if (fresh && !need) {
            pend[NSLOT - 1 - (nempty + __popc(md & lt))] = (uint8_t)eslot;
            pool.zs[eslot] = (uint32_t)DONE_NEW << 16;
          }
          nempty += __popc(md);
        }
        const unsigned mn = __ballot_sync(full, need);
        I3RC_STAT(W, ST_BIRTHS, __popc(mn));
        I3RC_STAT(W, ST_BIRTH_BATCH, mn != 0u);
        if (mn) {
          unsigned long long base = 0;
          const int leader = __ffs(mn) - 1;
          if (lane == leader) base = atomicAdd(p.nextPhoton, (unsigned long long)__popc(mn));
          base = __shfl_sync(full, base, leader);
          if (p.src.avail) {  // hand-filled arrays still being uploaded: wait until this warp's photons have arrived
            const unsigned long long want = min(base + (unsigned long long)__popc(mn), (unsigned long long)p.src.n);
            unsigned spins = 0;
            while (*(const volatile unsigned long long*)p.src.avail < want && ++spins < (1u << 24)) __nanosleep(256);
            if (spins >= (1u << 24)) I3RC_COUNT(E, CNT_BAD, 1);  // (4 s without the copy: give up rather than hang)
          }
          if (need) {
            const long long id = (long long)base + __popc(mn & lt);
            if (id < p.src.n) {
              xiTau = init_photon_state(p, E, id);
              E.eCell = ext_at(p, E.cx, E.cy, E.cz);
              pool.id[eslot] = (uint32_t)id;
              go = true;
            }
          }
          exhausted = (long long)base + __popc(mn) >= p.src.n;
        }
        // the next path segment becomes a task (MCRT:474-497)
        const unsigned mg = __ballot_sync(full, go);
        if (go) {
          LeTask t;
          t.xy = (uint32_t)E.cx | ((uint32_t)E.cy << 16);
          t.zdmc = (uint32_t)E.cz | ((uint32_t)MODE_PHOTON << 21);
          t.fx = E.fx;
          t.fy = E.fy;
          t.fz = E.fz;
          t.tauLimit = tau_of(xiTau);
          t.cw = __int_as_float(eslot);
          t.cfix = 0.0f;
          t.tauFree = xiTau;
          t.e0 = E.eCell;
          q[(tail + __popc(mg & lt)) & (QCAP - 1)] = t;
          pool.fx[eslot] = E.fx;
          pool.fy[eslot] = E.fy;
          pool.fz[eslot] = E.fz;
          pool.ux[eslot] = E.ux;
          pool.uy[eslot] = E.uy;
          pool.uz[eslot] = E.uz;
          pool.w[eslot] = E.w;
          pool.order[eslot] = E.order;
          pool.block[eslot] = E.rng.block;
        }
        tail += __popc(mg);
        stage = 0;
      }
      if (stage != 0) {  // suspended: park the photon in its slot, the rest of the batch state in the scratch
        I3RC_STAT(W, ST_SUSP, 1);
        if (has) {
          pool.xy[eslot] = (uint32_t)E.cx | ((uint32_t)E.cy << 16);
          pool.zs[eslot] = ((uint32_t)E.cz & 0xffffu) | ((uint32_t)(E.active ? DONE_INSIDE : DONE_NEW) << 16);
          pool.fx[eslot] = E.fx;
          pool.fy[eslot] = E.fy;
          pool.fz[eslot] = E.fz;
          pool.ux[eslot] = E.ux;
          pool.uy[eslot] = E.uy;
          pool.uz[eslot] = E.uz;
          pool.w[eslot] = E.w;
          pool.order[eslot] = E.order;
          pool.block[eslot] = E.rng.block;
        }
        uint32_t* sv = susp_of<NSLOT>(p, W, warp) + lane;
        sv[0 * 32] = __float_as_uint(a1);
        sv[1 * 32] = __float_as_uint(a2);
        sv[2 * 32] = __float_as_uint(a3);
        sv[3 * 32] = __float_as_uint(c2);
        sv[4 * 32] = __float_as_uint(c3);
        sv[5 * 32] = ((uint32_t)E.comp << 16) | ((uint32_t)E.pfi & 0xffffu);
        sv[6 * 32] = (uint32_t)eslot | (has ? 0x100u : 0u) | (alive ? 0x200u : 0u);
      }
    }
    __syncwarp();

    // ================= TRACE round =================
    // (pairs of steps: the two extinction registers of a ray swap roles on every step, see dda_step)
    // The round ends after STEPS crossings, or earlier when fewer than minRunning lanes still have a running ray: then
    // enough lanes wait for the (divergent, per-ray) bookkeeping below to make it worth its price.
    // (UNROLL pairs per loop body: unrolling keeps a ray's gathers in flight across pairs -- at a back-edge the compiler
    // waits for every outstanding load -- but measured slower, 2.05e8 vs 2.21e8 photons/s at 4 vs 1: instruction cache.)
    I3RC_STAT(W, ST_ROUNDS, 1);
    I3RC_STAT(W, ST_START_RUN, __popc(__ballot_sync(full, R.done == DONE_RUN)));
#pragma unroll 1
    for (int k = 0; k < STEPS / (2 * UNROLL); k++) {
      bool few = false;
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        I3RC_STAT(W, ST_PAIRS, 1);
        I3RC_STAT(W, ST_LANE_PAIRS, __popc(__ballot_sync(full, R.done == DONE_RUN)));
        if constexpr (SPLIT && I3RC_SPLIT_SINGLE_STEP) {
          // (the layer-table kernels carry the slab code in every step: one copy of the step in the loop instead of the
          //  two of a pair -- the extinction registers change places instead of roles -- halves a loop body that no longer
          //  fitted the instruction cache)
#pragma unroll 1
          for (int hstep = 0; hstep < 2; hstep++) {
            if (R.done == DONE_RUN) {
              dda_step<0>(p, R);
              if (R.done == DONE_RUN) {
                const float t = R.e0;
                R.e0 = R.e1;
                R.e1 = t;
              }
            }
          }
        } else if (R.done == DONE_RUN) {
          dda_step_pair(p, R);
        }
        few = __popc(__ballot_sync(full, R.done == DONE_RUN)) < minRunning;
        if (few) break;
      }
      if (few) break;
    }
    ray_after_steps(R);
    // close finished rays
    bool segEnd = false;
    Tally talR;
    if constexpr (TSM) talR.n = 0;
    if (R.done != DONE_RUN && R.done != DONE_IDLE) {
      uint32_t* rv = W.ray[0] + lane;  // what the ray is for waits in shared memory while it is traced
      if (R.mode == MODE_PHOTON) {
        const int done = R.done;
        R.slot = (int)rv[1 * 32];
        I3RC_COUNT(R, CNT_CROSS_PH, R.nsteps);
        if (FAST || p.useRayTracing) {
          const bool inside = done == DONE_INSIDE;
          if (!isinf(R.nax)) pool.fx[R.slot] = R.rx;  // (an axis the ray does not move along keeps its offset)
          if (!isinf(R.nay)) pool.fy[R.slot] = R.ry;
          if (!isinf(R.naz)) pool.fz[R.slot] = R.rz;
          pool.xy[R.slot] = (uint32_t)R.cntx | ((uint32_t)R.cnty << 16);
          pool.zs[R.slot] = (uint32_t)R.cntz | ((uint32_t)done << 16) | SLOT_RAW;
          if (inside) {
            pool.sp[R.slot] = R.sp;
            pool.spd[R.slot] = ray_stop_offset(R);
          }
        } else {  // maximum cross-section flight: the event point is already in the photon registers
          pool.fx[R.slot] = R.fx;
          pool.fy[R.slot] = R.fy;
          pool.fz[R.slot] = R.fz;
          pool.xy[R.slot] = (uint32_t)R.cx | ((uint32_t)R.cy << 16);
          pool.zs[R.slot] = ((uint32_t)R.cz & 0xffffu) | ((uint32_t)done << 16);
          pool.block[R.slot] = R.rng.block;
        }
        segEnd = true;
        R.done = DONE_IDLE;
      } else {
        I3RC_COUNT(R, CNT_CROSS_LE, R.nsteps);
        R.td = (int)(rv[0] & 31u);
        R.tcomp = (int)(rv[0] >> 8);
        R.tcw = __uint_as_float(rv[1 * 32]);
        R.tcfix = __uint_as_float(rv[2 * 32]);
        R.ttauFree = __uint_as_float(rv[3 * 32]);
        if (!finish_le_ray(p, R, talR)) R.done = DONE_IDLE;
      }
    }
    if constexpr (TSM) {  // the local-estimate contributions of this round
      const TallyLater& t = talR;
      warp_tally<TSM>(p, wt, t.n > 0, t.w0, t.o0, t.v0);
    }
    const unsigned ms = __ballot_sync(full, segEnd);
    if (segEnd) pend[npend + __popc(ms & lt)] = (uint8_t)W.ray[1][lane];
    npend += __popc(ms);
    I3RC_ASSERT(R, npend <= NSLOT && tail - head <= QCAP && tail - head >= 0);
    // idle lanes take the next tasks
    const bool idle = R.done == DONE_IDLE;
    const unsigned mi = __ballot_sync(full, idle);
    const int avail = tail - head;
    const int rank = __popc(mi & lt);
    if (idle && rank < avail) {
      // One path for both kinds of task: a path segment finds its direction in its photon slot, a local-estimate ray
      // in the copy of the direction table behind the slots; what the ray is for is parked in shared memory.
      const LeTask t = q[(head + rank) & (QCAP - 1)];
      const int mode = (int)((t.zdmc >> 21) & 7u);
      const int j = mode == MODE_PHOTON ? __float_as_int(t.cw) : NSLOT + (int)((t.zdmc >> 16) & 31u);
      I3RC_ASSERT(R, j >= 0 && j < NSLOT + MAX_DIRS && (int)(t.xy & 0xffffu) < p.nx && (int)(t.xy >> 16) < p.ny &&
                         (int)(t.zdmc & 0xffffu) < p.nz);
      const float ux = pool.ux[j], uy = pool.uy[j], uz = pool.uz[j];
      uint32_t* rv = W.ray[0] + lane;
      rv[0] = t.zdmc >> 16;  // d | mode << 5 | comp << 8
      rv[1 * 32] = __float_as_uint(t.cw);  // (a path segment: the slot number)
      rv[2 * 32] = __float_as_uint(t.cfix);
      rv[3 * 32] = __float_as_uint(t.tauFree);
      R.mode = mode;
      if (FAST || p.useRayTracing || mode != MODE_PHOTON) {
        start_ray_at(p, R, (int)(t.xy & 0xffffu), (int)(t.xy >> 16), (int)(t.zdmc & 0xffffu), t.fx, t.fy, t.fz, ux, uy, uz,
                     inv_abs(ux), inv_abs(uy), inv_abs(uz), t.tauLimit, t.e0);
      } else {  // maximum cross-section: the whole flight at once; its end is handled by the next round
        R.cx = (int)(t.xy & 0xffffu);
        R.cy = (int)(t.xy >> 16);
        R.cz = (int)(t.zdmc & 0xffffu);
        R.fx = t.fx;
        R.fy = t.fy;
        R.fz = t.fz;
        R.ux = ux;
        R.uy = uy;
        R.uz = uz;
        R.rng.init((uint64_t)(p.firstPhoton + (long long)pool.id[j]));
        R.rng.block = pool.block[j];
        R.nsteps = 0;
        max_cross_section_flight(p, R, t.tauFree);
      }
    }
    I3RC_STAT(W, ST_STARVED, max(__popc(mi) - avail, 0));
    I3RC_STAT(W, ST_RING_LEFT, max(avail - __popc(mi), 0));
    head += min(__popc(mi), avail);
    __syncwarp();
  }
  // flush the warp's counters
  __syncwarp();
  if (lane < CNT_N && W.cnt[lane]) atomicAdd(p.counters + lane, (unsigned long long)W.cnt[lane]);
  if (TSM) flush_staged_tallies<BLOCK>(p, s_dyn);
}

// ---- probes: accumulateExtinctionAlongPath for explicit rays (deterministic sub-path parity) ---------------
template <class P>
__global__ void k_trace_rays(const P p, int n, const float* __restrict__ pos, const float* __restrict__ dir,
                             const float* __restrict__ tauLimit, float* __restrict__ tauOut,
                             float* __restrict__ posOut, int* __restrict__ idxOut) {
  __shared__ uint32_t s_cnt[CNT_N];  // (event counters are atomics: shared memory, not a thread's local array)
  if (threadIdx.x < CNT_N) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  Lane L;
  L.cnt = s_cnt;
  L.mode = MODE_PHOTON;
  locate_abs(p.xe, p.xyRegular, p.nx, p.x0, p.xmax, p.dx, pos[3 * r], 1, &L.cx, &L.fx);
  locate_abs(p.ye, p.xyRegular, p.ny, p.y0, p.ymax, p.dy, pos[3 * r + 1], 1, &L.cy, &L.fy);
  locate_abs(p.ze, p.zRegular, p.nz, p.z0, p.zmax, p.dz, pos[3 * r + 2], 0, &L.cz, &L.fz);
  float dx = dir[3 * r], dy = dir[3 * r + 1], dz = dir[3 * r + 2];
  start_ray(p, L, dx, dy, dz, inv_abs(dx), inv_abs(dy), inv_abs(dz), tauLimit ? tauLimit[r] : INFINITY);
  while (L.done == DONE_RUN) {
    dda_step(p, L);
    ray_after_steps(L);
  }
  if (L.done == DONE_INSIDE) ray_stop_inside(p, L);
  tauOut[r] = L.done == DONE_BAD ? -2.0f : L.tau;
  ray_local(p, L, &L.fx, &L.fy, &L.fz);
  const int ix = ray_ix(p, L), iy = ray_iy(p, L), iz = ray_iz(p, L);
  if (posOut) {
    posOut[3 * r] = abs_x(p, ix, L.fx);
    posOut[3 * r + 1] = abs_y(p, iy, L.fy);
    posOut[3 * r + 2] = L.done == DONE_TOP ? p.zmax : (L.done == DONE_BOTTOM ? p.z0 : abs_z(p, iz, L.fz));
  }
  if (idxOut) {
    idxOut[3 * r] = ix + 1;
    idxOut[3 * r + 1] = iy + 1;
    idxOut[3 * r + 2] = iz + 1;
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

// Philox4x32-10 on the device for explicit (photon, block) counters: raw words and the deviates Rng::next4 makes of them
__global__ void k_probe_philox(uint32_t key0, uint32_t key1, int n, const unsigned long long* __restrict__ photon,
                               const uint32_t* __restrict__ block, uint32_t* __restrict__ raw, float* __restrict__ u) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng g;
  g.init(photon[i]);
  g.block = block[i];
  const u32x4 ctr = {g.id_lo, g.id_hi, g.block, 0u};
  const u32x4 r = philox4x32_10(ctr, key0, key1);
  raw[4 * i] = r.x, raw[4 * i + 1] = r.y, raw[4 * i + 2] = r.z, raw[4 * i + 3] = r.w;
  g.next4(key0, key1, u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
}
// next_direct (MCRT:2086-2113) exactly as the transport kernel runs it: deviates from the photon's Philox stream starting
// at the given block; returns the new direction and the number of blocks consumed
__global__ void k_probe_next_direct(uint32_t key0, uint32_t key1, int n, const unsigned long long* __restrict__ photon,
                                    const uint32_t* __restrict__ block, const float* __restrict__ S,
                                    const float* __restrict__ cosine, float* __restrict__ out,
                                    uint32_t* __restrict__ blocksUsed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ProblemDyn p;
  p.key0 = key0;
  p.key1 = key1;
  Lane L;
  L.rng.init(photon[i]);
  L.rng.block = block[i];
  L.ux = S[3 * i], L.uy = S[3 * i + 1], L.uz = S[3 * i + 2];
  next_direct(p, L, cosine[i]);
  out[3 * i] = L.ux, out[3 * i + 1] = L.uy, out[3 * i + 2] = L.uz;
  blocksUsed[i] = L.rng.block - block[i];
}

// ---- roofline ceilings measured on the spot ---------------------------------------------------
// Random 4-byte gathers over an array of nWords floats: every thread walks its own pseudo-random sequence of indices
// (a 32-bit LCG, so consecutive loads of a thread and the loads of neighbouring threads fall into unrelated sectors) and
// keeps UNR independent loads in flight.  Over an 8 MB array this is the L2 random-gather ceiling of the cell-crossing
// loop for an L2-resident extinction field; over an array far larger than L2 it is the HBM gather ceiling.
template <int UNR>
__global__ void __launch_bounds__(256) k_gather_bench(const float* __restrict__ a, uint32_t mask, int iters, float* __restrict__ sink) {
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.0f;
  for (int it = 0; it < iters; it++) {
    uint32_t idx[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      x = x * 1664525u + 1013904223u;
      idx[u] = (x >> 7) & mask;
    }
#pragma unroll
    for (int u = 0; u < UNR; u++) acc += __ldg(a + idx[u]);
  }
  if (acc == -1.0f) sink[0] = acc;  // (never true for the zero-filled array: keeps the loads alive)
}
// Issue-slot ceiling: independent FMAs only, every warp scheduler can issue one warp instruction per cycle
__global__ void __launch_bounds__(256) k_issue_bench(int iters, float* __restrict__ sink) {
  float a0 = threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 1.0000001f, c = 1e-9f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      a0 = fmaf(a0, m, c), a1 = fmaf(a1, m, c), a2 = fmaf(a2, m, c), a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c), a5 = fmaf(a5, m, c), a6 = fmaf(a6, m, c), a7 = fmaf(a7, m, c);
    }
  }
  const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == -1.0f) sink[0] = s;
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
// One component's extinction replaced by a horizontally uniform profile (a k-distribution term of a gas component).
// The un-normalised extinction of every component is kept on the device from the first swap on (k_recover_extinctions:
// from the cumulative fractions, once), so that swapping band after band does not accumulate rounding: every swap
// rebuilds totals and fractions from those, like getOpticalPropertiesByComponent does (Code/opticalProperties.f95:524-537).
__global__ void k_recover_extinctions(size_t ncell, int nc, const float* __restrict__ totalExt, const float* __restrict__ cumExt,
                                      float* __restrict__ raw) {
  const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  const float tot = totalExt[cell];
  float prev = 0.0f;
  for (int c = 0; c < nc; c++) {
    float frac = cumExt[(size_t)c * ncell + cell];
    if (c == nc - 1 && tot > F_TINY) frac = 1.0f;  // (undo the 1 + epsilon nudge)
    raw[(size_t)c * ncell + cell] = tot > F_TINY ? (frac - prev) * tot : frac - prev;
    prev = frac;
  }
}
__global__ void k_replace_profile(int nx, int ny, int nz, int nc, int comp, const float* __restrict__ profile,
                                  float* __restrict__ raw, float* __restrict__ totalExt, float* __restrict__ cumExt) {
  const size_t ncell = (size_t)nx * ny * nz;
  const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  const int iz = (int)(cell / ((size_t)nx * ny));
  raw[(size_t)comp * ncell + cell] = profile[iz];
  float run = 0.0f;
  for (int c = 0; c < nc; c++) {
    const float e = raw[(size_t)c * ncell + cell];
    run = (c == 0) ? e : e + run;
    cumExt[(size_t)c * ncell + cell] = run;
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

// totalExt[z][y][x] -> z-fastest copy out[x][y][z] (tiled through shared memory; one x-z plane per blockIdx.z)
__global__ void k_transpose_zfast(int nx, int ny, int nz, const float* __restrict__ in, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int iy = blockIdx.z;
  const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const int ix = x0 + threadIdx.x, iz = z0 + k;
    if (ix < nx && iz < nz) tile[k][threadIdx.x] = in[((size_t)iz * ny + iy) * nx + ix];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const int ix = x0 + k, iz = z0 + threadIdx.x;
    if (ix < nx && iz < nz) out[((size_t)ix * ny + iy) * nz + iz] = tile[threadIdx.x][k];
  }
}

// ---- empty-space codes (transport.cuh, JUMP_MIN / JUMP_MAX) -----------------------------------------------------------
// D[cell] = Chebyshev distance (in cells; periodic in x and y, nothing beyond the top and the bottom) to the nearest cell
// with extinction, grown one shell per launch: 0 = the cell has extinction, 255 = not reached yet.  Layout of the gather
// field (z fastest).
__global__ void k_empty_init(const float* __restrict__ ext, size_t n, uint8_t* __restrict__ D) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) D[i] = ext[i] > 0.0f ? 0 : 255;
}
__global__ void k_empty_grow(int nx, int ny, int nz, int k, uint8_t* D) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nx * ny * nz || D[i] != 255) return;
  const int iz = (int)(i % nz), iy = (int)((i / nz) % ny), ix = (int)(i / ((size_t)nz * ny));
  bool hit = false;
  for (int dx = -1; dx <= 1 && !hit; dx++) {
    const int jx = (ix + dx + nx) % nx;
    for (int dy = -1; dy <= 1 && !hit; dy++) {
      const int jy = (iy + dy + ny) % ny;
      for (int dz = -1; dz <= 1; dz++) {
        const int jz = iz + dz;
        if (jz < 0 || jz >= nz) continue;
        if (D[((size_t)jx * ny + jy) * nz + jz] == k - 1) {  // (shell k-1 is complete; this launch only writes k)
          hit = true;
          break;
        }
      }
    }
  }
  if (hit) D[i] = (uint8_t)k;
}
// coded copy of the field: an empty cell at distance D >= jumpMin + 2 holds -(min(D - 2, JUMP_MAX)); counts the coded cells
__global__ void k_empty_code(const float* __restrict__ ext, const uint8_t* __restrict__ D, size_t n, int jumpMin,
                             float* __restrict__ coded, unsigned long long* __restrict__ count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e = ext[i];
  const int d = D[i];
  if (d >= jumpMin + 2) {
    e = -(float)min(d - 2, JUMP_MAX);
    atomicAdd(count, 1ull);
  }
  coded[i] = e;
}

// leLB[d][cell] and leUB[d][cell] for the radiance directions (Problem::leLB, leUB): one thread per cell and direction; only cells a ray can start
// from matter (cells with extinction, and the bottom layer where the surface reflects), the others get 0
__global__ void k_le_path_bounds(int nx, int ny, int nz, float dx, float dy, float dz, const float* __restrict__ ext,
                                 const float* __restrict__ dirs, int nDir, int nLayers, float* __restrict__ lower,
                                 float* __restrict__ upper) {
  // Direction-fastest: lower[cell * nDir + d].  An event asks for the bounds of ONE cell towards all directions, one after
  // the other: they share one or two 32-byte sectors instead of lying a whole field apart (twelve slanted directions on
  // 512x512x256: twelve dependent HBM round trips per event became one).
  const size_t ncell = (size_t)nx * ny * nz;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  const int ix = (int)(i % nx), iy = (int)((i / nx) % ny), iz = (int)(i / ((size_t)nx * ny));
  const bool ask = ext[i] > 0.0f || iz == 0;
  for (int d = 0; d < nDir; d++) {
    float lo = 0.0f, hi = INFINITY;
    if (ask)
      le_path_bounds(ext, nx, ny, nz, dx, dy, dz, dirs[d * DIR_STRIDE], dirs[d * DIR_STRIDE + 1], dirs[d * DIR_STRIDE + 2], ix, iy, iz,
                     nLayers, LE_LB_ENOUGH, &lo, &hi);
    lower[i * nDir + d] = lo;
    if (upper) upper[i * nDir + d] = hi;
  }
}

// colTau[k][col] = sum over layers j >= k of totalExt[j][col] * (ze[j+1] - ze[j]), k = 0 .. nz (Problem::colTau)
__global__ void k_column_suffix(int nz, size_t ncol, const float* __restrict__ ext, const float* __restrict__ ze,
                                float* __restrict__ S) {
  const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  float run = 0.0f;
  S[(size_t)nz * ncol + col] = 0.0f;
  for (int k = nz - 1; k >= 0; k--) {
    run = fmaf(ext[(size_t)k * ncol + col], ze[k + 1] - ze[k], run);
    S[(size_t)k * ncol + col] = run;
  }
}

// per layer: are all cells of the layer equal?  out[iz] = {min bits, max bits} (extinctions are >= 0: bits order = value order)
__global__ void k_layer_minmax(const float* __restrict__ in, size_t ncol, unsigned int* __restrict__ out) {
  __shared__ unsigned int lo[256], hi[256];
  const float* row = in + (size_t)blockIdx.x * ncol;
  unsigned int a = 0xffffffffu, b = 0u;
  for (size_t i = threadIdx.x; i < ncol; i += blockDim.x) {
    const unsigned int v = __float_as_uint(row[i]);
    a = min(a, v);
    b = max(b, v);
  }
  lo[threadIdx.x] = a;
  hi[threadIdx.x] = b;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      lo[threadIdx.x] = min(lo[threadIdx.x], lo[threadIdx.x + o]);
      hi[threadIdx.x] = max(hi[threadIdx.x], hi[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = lo[0];
    out[2 * blockIdx.x + 1] = hi[0];
  }
}
// the layers that vary horizontally, z-fastest: out[(ix*ny + iy)*nzc + k] = in[layer[k]][iy][ix]
__global__ void k_compact_zfast(int nx, int ny, int nzc, const int* __restrict__ layer, const float* __restrict__ in,
                                float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int iy = blockIdx.z;
  const int x0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const int ix = x0 + threadIdx.x, kk = k0 + k;
    if (ix < nx && kk < nzc) tile[k][threadIdx.x] = in[((size_t)layer[kk] * ny + iy) * nx + ix];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const int ix = x0 + k, kk = k0 + threadIdx.x;
    if (ix < nx && kk < nzc) out[((size_t)ix * ny + iy) * nzc + kk] = tile[threadIdx.x][k];
  }
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
// fluxAbsorbed[col] = sum over layers of the raw volume absorption (the same increments, MCRT:644-647; Problem::deriveAbs)
__global__ void k_abs_from_volume(int nz, size_t ncol, const float* __restrict__ volAbs, float* __restrict__ fluxAbs) {
  const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  float s = 0.0f;
  for (int k = 0; k < nz; k++) s += volAbs[(size_t)k * ncol + col];
  fluxAbs[col] = s;
}
// main[i] += sum over the K-1 copies of copy[k][i] (in float64), copies zeroed for the next launch (Problem::rep)
__global__ void k_fold_replicas(float* __restrict__ main, float* __restrict__ rep, size_t n, int copies, size_t stride) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = (double)main[i];
  for (int k = 0; k < copies; k++) {
    s += (double)rep[(size_t)k * stride + i];
    rep[(size_t)k * stride + i] = 0.0f;
  }
  main[i] = (float)s;
}
// A batch traced in pieces: acc += tally; the tally restarts from zero, or -- after the last piece -- gets the total
__global__ void k_fold_tally(float* __restrict__ tally, double* __restrict__ acc, size_t n, int last) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double s = acc[i] + (double)tally[i];
    acc[i] = s;
    tally[i] = last ? (float)s : 0.0f;
  }
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
