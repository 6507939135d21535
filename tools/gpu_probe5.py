#!/usr/bin/env python
"""One workload, several tunings (development aid): gpu_probe5.py <workload> <photons> '<tune dict>' ..."""
import sys

sys.path.insert(0, ".")
from tools.gpu_probe2 import run

if __name__ == "__main__":
    wl, nph = sys.argv[1], int(sys.argv[2])
    for t in sys.argv[3:] or ["{}"]:
        run(wl, nph, 1, eval(t))
