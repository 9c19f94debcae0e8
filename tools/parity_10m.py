#!/usr/bin/env python3
"""Headline parity run: validity of 10 M Panda states (config C2 scene) on the GPU vs the oracle, bit for bit.

Both sides read the SAME distance fields (the reference's own propagation result, injected through
b200sv_field_set_d2), so this isolates K1 + K2 from the EDT's documented differences; the field difference is
reported separately.  Mismatches are bucketed by the oracle's clearance as north_star asks.
Writes one JSON line; run on the GPU box:  python tools/parity_10m.py [n_states] > gpurun_out/parity.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("ORACLE_MARCH", "native")
import oracle  # noqa: E402
oracle.build(force=True)
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402
from util import config_pair  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
name = sys.argv[2] if len(sys.argv) > 2 else "C2"
eng, env, cloud, _ = config_pair(name, n_states=1)
env.world.add_points(cloud)
eng.field_add_points(cloud)
g, o = eng.field_get_d2(), env.world.d2()
out = {"workload": name, "states": n, "edt_voxels": int(g.size), "edt_voxels_gpu_below_reference": int((g < o).sum()),
       "edt_voxels_gpu_above_reference": int((g > o).sum())}
eng.field_set_d2(o, mb.FIELD_WORLD)
eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
lo, hi = eng.group_bounds()
mism, near, rechecked, valid_total = 0, 0, 0, 0
t_gpu = t_cpu = 0.0
chunk = 1_000_000
for c0 in range(0, n, chunk):
    m = min(chunk, n - c0)
    q = workloads.random_states(1000 + c0 // chunk, m, lo, hi)
    t = time.perf_counter()
    valid = eng.check_batch(q, mb.CHECK_ALL)
    t_gpu += time.perf_counter() - t
    rechecked += eng.stats()["n_rechecked"]
    t = time.perf_counter()
    col, cores = env.check(q, 7)
    t_cpu += time.perf_counter() - t
    bad = np.nonzero(valid == col.astype(bool))[0]
    valid_total += int(valid.sum())
    for i in bad:
        mism += 1
        if abs(env.margin(q[i], 7)) < 1e-4:
            near += 1
out.update({"validity_mismatches": mism, "mismatches_with_reference_clearance_below_1e-4m": near,
            "mismatches_elsewhere": mism - near, "valid_states": valid_total, "fp64_rechecked_states": int(rechecked),
            "gpu_seconds_incl_copies": round(t_gpu, 3), "oracle_seconds": round(t_cpu, 3), "oracle_threads": cores})
print(json.dumps(out))
