#!/usr/bin/env python3
"""Measure the FP32 pass's pose error against the FP64 pass (what the fp32_eps budget must cover)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb
from moveit2_b200 import workloads
for robot, group in (("panda", "panda_arm"), ("dualarm", "whole_body")):
    eng = mb.StateValidityEngine.for_robot(robot, group, size=(2, 2, 2), origin=(0, 0, 0.5), resolution=0.01, max_distance=0.25)
    lo, hi = eng.group_bounds()
    worst_t, worst_r = 0.0, 0.0
    for seed in range(10):
        q = workloads.random_states(100 + seed, 100000, lo, hi)
        a, b = eng.fk_batch(q, 32), eng.fk_batch(q, 64)
        d = np.abs(a - b)
        worst_t = max(worst_t, d[:, :, :3, 3].max())
        worst_r = max(worst_r, d[:, :, :3, :3].max())
    sp = eng.model_arrays()["sphere_xyzr"]
    reach = np.abs(sp[:, :3]).sum(axis=1).max()
    print("%s/%s: max |dt| = %.3e m, max |dR| = %.3e, max sphere offset L1 = %.3f m => centre error <= %.3e m" %
          (robot, group, worst_t, worst_r, reach, worst_t + worst_r * reach))
