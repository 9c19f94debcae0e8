#!/usr/bin/env python3
"""End-to-end rate of the host-buffer entry points on C2 (pinned buffers): f64 and f32 states.  B200SV_CHUNK sets the
copy / kernel pipeline's chunk (states)."""
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
n = 1_000_000
q = workloads.random_states(2, n, lo, hi)
qp = torch.from_numpy(q).pin_memory()
qf = torch.from_numpy(q.astype(np.float32)).pin_memory()
bm = torch.zeros((n + 7) // 8, dtype=torch.uint8).pin_memory()
for name, fn, ptr in (("f64", eng.check_batch_ptr, qp.data_ptr()), ("f32", eng.check_batch_f32_ptr, qf.data_ptr())):
    for _ in range(5):
        fn(ptr, n, mb.CHECK_ALL, bm.data_ptr())
    t0 = time.perf_counter()
    for _ in range(50):
        fn(ptr, n, mb.CHECK_ALL, bm.data_ptr())
    dt = (time.perf_counter() - t0) / 50
    print("%s chunk=%s: %.4f ms per 1M states = %.3g states/s (kernels inside %.3f ms)" %
          (name, os.environ.get("B200SV_CHUNK", "default"), dt * 1e3, n / dt, eng.stats()["check_ms"]))
