#!/usr/bin/env python3
"""Generates tests/golden/df_golden.npz from the reference's checked-in distance-field fixtures.

Run in the authoring container only (needs /root/reference).  The .df files hold OCCUPANCY only
(propagation_distance_field.cpp:635-673); what is committed here is the decoded occupied-voxel set of each
(bit-packed), i.e. derived golden data, not a copy of the fixture files:
  test_small.df  10^3 @ 0.1 m:  the three voxels of POINT1-3 (test_distance_field.cpp:58-60, 946-950)
  test_big.df    150x150x200 @ 0.02 m: sphere r=0.5 at (0.5,0.5,0.5) via addShapeToField (:957-963)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.df import read_df_occupancy  # noqa: E402

REF = "/root/reference/moveit_core"
out = {}
for name in ("test_small", "test_big"):
    hdr, n, occ = read_df_occupancy(open(os.path.join(REF, name + ".df"), "rb").read())
    grid = np.zeros(n, np.uint8)
    grid[occ[:, 0], occ[:, 1], occ[:, 2]] = 1
    out[name + "_header"] = np.array([hdr[k] for k in ("resolution", "size_x", "size_y", "size_z", "origin_x",
                                                       "origin_y", "origin_z")])
    out[name + "_dims"] = np.array(n, np.int32)
    out[name + "_occ_bits"] = np.packbits(grid.reshape(-1))
    print(name, hdr, n, len(occ))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "df_golden.npz"), **out)
