"""The numpy Philox used as the checker of the device stream, against the Random123 known-answer vectors."""
import numpy as np

from tests.philox_numpy import philox4x32_10, photon_block


def test_numpy_philox_known_answers():
    assert philox4x32_10([[0, 0, 0, 0]], (0, 0)).tolist() == [[0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]]
    assert philox4x32_10([[0xFFFFFFFF] * 4], (0xFFFFFFFF, 0xFFFFFFFF)).tolist() == [[0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]]
    assert philox4x32_10([[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], (0xA4093822, 0x299F31D0)).tolist() == [
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]]


def test_numpy_philox_matches_the_cpu_build_of_the_device_header():
    from tests.hostsim.binding import philox
    rng = np.random.default_rng(5)
    for _ in range(20):
        ph, bl = int(rng.integers(0, 2**40)), int(rng.integers(0, 1000))
        key = (int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32)))
        want = philox((ph & 0xFFFFFFFF, ph >> 32, bl, 0), key)
        assert photon_block(key, [ph], bl)[0].tolist() == want
