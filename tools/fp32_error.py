#!/usr/bin/env python3
"""The FP32 budget held to account: posed sphere centres of the FAST kernel's arithmetic (b200sv_debug_fast_centres: the
kernel's own FK function) against the exact FP64 centres, for the bench configs; prints the worst error next to the budget
derived from the model (b200sv_get_fp32_budget).  python tools/fp32_error.py [n_states]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
out = {}
for name in ("C1", "C2", "C3", "C3D", "C3W"):
    cfg = workloads.CONFIGS[name]
    eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                           resolution=cfg["resolution"], max_distance=cfg["max_distance"])
    lo, hi = eng.group_bounds()
    lo, hi = np.maximum(lo, -2.0 * np.pi), np.minimum(hi, 2.0 * np.pi)
    q = workloads.random_states(555, n, lo, hi)
    # the corners of the joint box are where |q| and the reach are largest
    q[: 2 ** min(len(lo), 8)] = np.array([[hi[k] if (i >> (k % 8)) & 1 else lo[k] for k in range(len(lo))]
                                          for i in range(2 ** min(len(lo), 8))])
    fast = eng.debug_fast_centres(q)
    exact = np.stack([eng.sphere_centers(q[i]) for i in range(len(q))])
    err = np.abs(fast - exact).max(axis=2)       # per sphere, the largest axis error (the budget is per axis)
    used, derived = eng.fp32_budget()
    out[name] = {"robot": cfg["robot"], "group": cfg["group"], "states": int(n), "max_error_m": float(err.max()),
                 "p999_error_m": float(np.quantile(err, 0.999)), "budget_used_m": used, "budget_derived_m": derived,
                 "worst_over_budget": float(err.max() / used)}
    eng.close()
print(json.dumps(out, indent=1))
