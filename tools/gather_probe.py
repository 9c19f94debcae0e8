#!/usr/bin/env python3
"""The gather-roofline microbenchmark alone, on the C2 field footprint (16 MB interleaved uint8 field), for an ncu capture:
  ncu --set full -k regex:gatherKernel -c 2 -o gpurun_out/gather python tools/gather_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

cfg = workloads.CONFIGS["C2"]
eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                       resolution=cfg["resolution"], max_distance=cfg["max_distance"])
eng.field_add_points(workloads.make_cloud("C2"))
print(json.dumps({"gather_gbs_32B_sectors": eng.gather_roofline(1 << 28, 3), "footprint_bytes": 2 * 200 ** 3}))
