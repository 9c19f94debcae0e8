#!/usr/bin/env python3
"""Timeline of the host entry points' copy / kernel pipeline on C2 (B200SV_TRACE=1 makes the library print, per chunk, when
its copy landed and when its kernel finished): e2e_trace.py [f64|f32]"""
import os, sys
os.environ["B200SV_TRACE"] = "1"
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
n = 1_000_000
q = workloads.random_states(2, n, lo, hi)
qp = torch.from_numpy(q).pin_memory()
qf = torch.from_numpy(q.astype(np.float32)).pin_memory()
bm = torch.zeros((n + 7) // 8, dtype=torch.uint8).pin_memory()
for name, fn, ptr in (("f64", eng.check_batch_ptr, qp.data_ptr()), ("f32", eng.check_batch_f32_ptr, qf.data_ptr())):
    for _ in range(4):
        fn(ptr, n, mb.CHECK_ALL, bm.data_ptr())
