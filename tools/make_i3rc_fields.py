#!/usr/bin/env python
"""Pack the I3RC Phase-1 input fields shipped with the reference into one small fixture.

Reads (in THIS container only) the ASCII data files under /root/reference/I3RC-Examples/Data and writes
i3rc_monte_carlo_model_b200/data/i3rc_fields.npz, which travels with the repo (the GPU box has no
/root/reference).  These are input DATA (optical depth / thickness maps and one tabulated phase
function), not reference source code.  Formats follow the read statements of
I3RC-Examples/i3rcLandsatCloud.f95:70-80 ('(128f7.2)', rows = y) and i3rcRadarCloud.f95:66-114
('(640f8.3)', rows top -> bottom; C.1_PF = 1801 "angle value" lines; C.1_leg_coef = 300 lines).
"""
import os
import sys

import numpy as np

SRC = "/root/reference/I3RC-Examples/Data"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "i3rc_monte_carlo_model_b200", "data", "i3rc_fields.npz")


def fixed(path, width, ncol):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line.strip():
                continue
            rows.append([float(line[i * width:(i + 1) * width]) for i in range(ncol)])
    return np.array(rows, dtype=np.float32)


def main():
    tau = fixed(os.path.join(SRC, "scene43.tau.128x128"), 7, 128)   # [y][x]
    dz = fixed(os.path.join(SRC, "scene43.dz.128x128"), 7, 128)     # km
    radar = fixed(os.path.join(SRC, "mmcr_tau_32km_020898"), 8, 640)  # [row top->bottom][x]
    pf = np.loadtxt(os.path.join(SRC, "C.1_PF"), dtype=np.float64)
    leg = np.loadtxt(os.path.join(SRC, "C.1_leg_coef"), dtype=np.float64)
    assert tau.shape == (128, 128) and dz.shape == (128, 128) and radar.shape == (54, 640), (tau.shape, dz.shape, radar.shape)
    assert pf.shape == (1801, 2)
    np.savez_compressed(DST, landsat_tau=tau, landsat_dz_km=dz, radar_tau=radar,
                        c1_angle_deg=pf[:, 0].astype(np.float32), c1_value=pf[:, 1].astype(np.float32),
                        c1_leg_coef=np.ravel(leg).astype(np.float32))
    print("wrote", DST, os.path.getsize(DST), "bytes")
    print("landsat: mean tau %.3f max %.2f cloud fraction %.3f" % (tau.mean(), tau.max(), (tau > 0).mean()))
    print("radar: mean column tau %.2f max %.2f" % (radar.sum(0).mean(), radar.sum(0).max()))


if __name__ == "__main__":
    sys.exit(main())
