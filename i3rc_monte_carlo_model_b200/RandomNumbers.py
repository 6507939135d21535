"""``randomNumberSequence`` at the boundary (Code/RandomNumbersForMC.f95:99-239).

The CUDA path does not run MT19937: the seed vector the driver passes to ``new_RandomNumberSequence``
((/ iseed, batch /) in monteCarloDriver.f95:277, (/ batch, iseed /) in planeParallel.f95:207) becomes the
key of the counter-based Philox4x32-10 streams, one stream per photon.  The object therefore only
carries the seed vector.
"""
from __future__ import annotations

import numpy as np


class randomNumberSequence:
    def __init__(self, seed):
        self.seed = np.atleast_1d(np.asarray(seed, dtype=np.int32)).copy()


def new_RandomNumberSequence(seed):
    return randomNumberSequence(seed)


def finalize_RandomNumberSequence(twister):
    twister.seed = np.zeros(0, np.int32)
