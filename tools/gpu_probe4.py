#!/usr/bin/env python
"""Throughput vs batch size (tail effect of the persistent kernel) and vs tuning (development aid)."""
import sys

sys.path.insert(0, ".")
from tools.gpu_probe2 import run

if __name__ == "__main__":
    tunes = [eval(a) for a in sys.argv[1:]] or [{}]
    for tune in tunes:
        for nph in (1_000_000, 4_000_000, 16_000_000, 64_000_000):
            print(nph, end=" ")
            run("landsat", nph, 1, tune)
