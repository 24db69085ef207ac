#!/usr/bin/env python
"""Regenerate gipuma_b200/data/dtu_calib_P.npy from the reference's calibration fixtures.

The only data the reference ships is data/dtu/calib/*.P: 3x4 DTU projection matrices for the 64
camera positions.  This script parses the 64 `rect_0NN_3_r5000.png.P`
files (text, three rows of four numbers, some with CRLF) into one float64 array [64, 3, 4] so the
synthetic benchmark scenes use real DTU camera geometry on the GPU box, where /root/reference
does not exist.  Run here (needs /root/reference):  python tools/make_dtu_fixture.py
"""
import os
import sys
import numpy as np

ref = os.environ.get("GIPUMA_REFERENCE", "/root/reference")
calib = os.path.join(ref, "data", "dtu", "calib")
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gipuma_b200", "data", "dtu_calib_P.npy")

Ps = []
for k in range(1, 65):
    path = os.path.join(calib, "rect_%03d_3_r5000.png.P" % k)
    with open(path) as fh:
        vals = [float(t) for t in fh.read().split()]
    assert len(vals) == 12, path
    Ps.append(np.array(vals, dtype=np.float64).reshape(3, 4))
Ps = np.stack(Ps)
np.save(out, Ps)
print("wrote", os.path.normpath(out), Ps.shape, file=sys.stderr)
