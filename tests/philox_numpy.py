"""An independent numpy implementation of Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
1, 2, 3", SC'11), vectorised over counters.  Test infrastructure: the device stream (csrc/philox.cuh) is checked against
it, and it is itself checked against the Random123 known-answer vectors (tests/test_philox_numpy.py)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: [n, 4] uint32 counters, key: (k0, k1) -> [n, 4] uint32"""
    c = np.array(ctr, dtype=np.uint64).reshape(-1, 4).copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n = np.empty_like(c)
        n[:, 0] = hi1 ^ c[:, 1] ^ np.uint64(k0)
        n[:, 1] = lo1
        n[:, 2] = hi0 ^ c[:, 3] ^ np.uint64(k1)
        n[:, 3] = lo0
        c = n
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def photon_block(key, photon, block):
    """The four words of block `block` of photon `photon`'s stream (counter layout of csrc/philox.cuh)."""
    photon = np.atleast_1d(np.asarray(photon, np.uint64))
    block = np.broadcast_to(np.asarray(block, np.uint64), photon.shape)
    ctr = np.stack([photon & MASK, photon >> np.uint64(32), block, np.zeros_like(photon)], axis=1)
    return philox4x32_10(ctr, key)


def u01(words):
    """float32 deviate on [0,1] (closed, like genrand_real1 cast to REAL, RandomNumbersForMC.f95:275-299)"""
    return (words.astype(np.float32) * np.float32(2.3283064365386963e-10)).astype(np.float32)
