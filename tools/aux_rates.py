#!/usr/bin/env python3
"""Throughput of the auxiliary batched entry points on config C2 (kernel time from b200sv_get_stats where the entry
point records it, wall clock incl. copies otherwise).  python tools/aux_rates.py > gpurun_out/aux_rates.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

cfg = workloads.CONFIGS["C2"]
eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                       resolution=cfg["resolution"], max_distance=cfg["max_distance"])
eng.field_add_points(workloads.make_cloud("C2"))
lo, hi = eng.group_bounds()
q = workloads.random_states(77, 200_000, lo, hi)
out = {"workload": "C2 scene, Panda panda_arm, %d states" % len(q)}
for _ in range(2):
    eng.collision_gradients(q)
out["collision_gradients_states_per_s_kernel"] = len(q) / (eng.stats()["check_ms"] * 1e-3)
for _ in range(2):
    t = time.perf_counter()
    eng.check_contacts(q, max_contacts=8, max_contacts_per_pair=2)
    dt = time.perf_counter() - t
out["check_contacts_states_per_s_wall"] = len(q) / dt
L = 0.01 * float(np.sum(hi - lo))
a, b = q[:100_000], q[100_000:]
for _ in range(2):
    valid, first, n = eng.check_motions(a, b, L)
out["check_motions_edges"] = len(a)
out["check_motions_states_expanded"] = n
out["check_motions_states_per_s_kernels"] = n / (eng.stats()["check_ms"] * 1e-3)
out["check_motions_edges_per_s_kernels"] = len(a) / (eng.stats()["check_ms"] * 1e-3)
print(json.dumps(out))
