#!/usr/bin/env python3
"""Wall-clock latency of b200sv_check_batch (host buffers in, bitmap out) for small batches on config C1 -- what a
single-state CollisionEnv::checkCollision call through the plugin pays.  python tools/latency.py > gpurun_out/latency.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

cfg = workloads.CONFIGS["C1"]
eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                       resolution=cfg["resolution"], max_distance=cfg["max_distance"])
eng.field_add_points(workloads.make_cloud("C1"))
lo, hi = eng.group_bounds()
out = {"workload": "C1 scene, Panda panda_arm, b200sv_check_batch wall clock per call (median of 200), pageable host buffers"}
for n in (1, 32, 1000, 10000, 100000):
    q = workloads.random_states(3, n, lo, hi)
    for _ in range(20):
        eng.check_batch(q, mb.CHECK_ALL)
    ts = []
    for _ in range(200):
        t = time.perf_counter()
        eng.check_batch(q, mb.CHECK_ALL)
        ts.append(time.perf_counter() - t)
    out["n=%d" % n] = {"median_us": 1e6 * float(np.median(ts)), "p90_us": 1e6 * float(np.percentile(ts, 90)),
                       "states_per_s": n / float(np.median(ts))}
print(json.dumps(out))
