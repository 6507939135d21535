// Philox4x32-10 counter-based generator (Salmon et al., SC'11), one stream per photon.
// Replaces the shared sequential MT19937 stream of Code/RandomNumbersForMC.f95 on the device:
//   key     = (seed[0], seed[1])  -- the vector the driver gives new_RandomNumberSequence
//                                    ((/ iseed, batch /), monteCarloDriver.f95:277)
//   counter = (photon id lo, photon id hi, draw block, stream tag)
// so any (batch, photon) is reproducible independently of the GPU count or the launch shape.
#pragma once
#include <stdint.h>

#ifndef I3RC_HD
#ifdef __CUDACC__
#define I3RC_HD __host__ __device__ __forceinline__
#else
#define I3RC_HD inline
#endif
#endif

namespace i3rc {

struct u32x4 {
  uint32_t x, y, z, w;
};

I3RC_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

#if defined(__CUDACC__) && defined(I3RC_PHILOX_NOINLINE)  // (experiment: one copy of the ten rounds in the kernel instead of four)
__host__ __device__ __noinline__
#else
I3RC_HD
#endif
u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    u32x4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// uniform on [0,1] with the reference's closed-interval convention (genrand_real1 cast to REAL,
// RandomNumbersForMC.f95:275-299): exactly 0 and exactly 1 are possible and guarded by the callers.
I3RC_HD float u01(uint32_t x) { return (float)x * 2.3283064365386963e-10f; }

// Per-photon stream, consumed in whole BLOCKS of four deviates.  The transport code asks for a fresh block at
// fixed points of a photon's life (birth, every event, every pair of local-estimate directions, every pair of
// rejection rounds of the direction change), so that all lanes of a warp that are in the same phase generate their
// blocks together instead of refilling a per-lane buffer at data-dependent times.
struct Rng {
  uint32_t id_lo, id_hi, block;
  I3RC_HD void init(uint64_t photon) {
    id_lo = (uint32_t)photon;
    id_hi = (uint32_t)(photon >> 32);
    block = 0;
  }
  I3RC_HD void next4(uint32_t k0, uint32_t k1, float& a, float& b, float& c, float& d) {
    u32x4 ctr = {id_lo, id_hi, block, 0u};
    u32x4 r = philox4x32_10(ctr, k0, k1);
    block++;
    a = u01(r.x);
    b = u01(r.y);
    c = u01(r.z);
    d = u01(r.w);
  }
};

}  // namespace i3rc
