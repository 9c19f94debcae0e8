#!/usr/bin/env python3
"""e2e rate of b200sv_check_motions on C2 (100 000 random edges, pinned host ends); B200SV_TRACE=1 prints the library's own
timeline of a call."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb
from moveit2_b200 import workloads
cfg = workloads.CONFIGS["C2"]
eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                       resolution=cfg["resolution"], max_distance=cfg["max_distance"])
eng.field_add_points(workloads.make_cloud("C2"))
lo, hi = eng.group_bounds()
ne = 100_000
q = torch.from_numpy(workloads.random_states(2, 2 * ne, lo, hi)).pin_memory().numpy()
ea, eb = q[:ne], q[ne:]
seg = 0.01 * float(np.sum(hi - lo))
for _ in range(3):
    eng.check_motions(ea, eb, seg)
t0 = time.perf_counter()
for _ in range(5):
    ok, first, n_exp = eng.check_motions(ea, eb, seg)
dt = (time.perf_counter() - t0) / 5
print("edges: %.3f ms per %d edges (%d states) = %.3g states/s; valid edges %d; device span %.3f ms" %
      (dt * 1e3, ne, n_exp, n_exp / dt, int(ok.sum()), eng.stats()["check_ms"]))
