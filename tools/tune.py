#!/usr/bin/env python3
"""Sweep the check kernel's launch knobs (states per warp per FK pass, CTAs per SM) on the C2 workload."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
cfg = workloads.CONFIGS[name]
eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                       resolution=cfg["resolution"], max_distance=cfg["max_distance"])
eng.field_add_points(workloads.make_cloud(name))
lo, hi = eng.group_bounds()
q = torch.from_numpy(workloads.random_states(cfg["states"]["seed"], n, lo, hi)).cuda()
bm = torch.zeros(4 * ((n + 31) // 32), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()


def run(flags, reps=10):
    for _ in range(3):
        eng.check_batch_device(q.data_ptr(), n, flags, bm.data_ptr(), s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        eng.check_batch_device(q.data_ptr(), n, flags, bm.data_ptr(), s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ref = None
for warps in (16, 12, 8, 6, 4):
    for blocks in (1, 2, 3):
        try:
            eng.set_tuning(0.0, warps, blocks)
        except mb.B200svError as e:
            print("warps %2d blocks %d: %s" % (warps, blocks, e))
            continue
        ms = run(mb.CHECK_ALL)
        out = bm.cpu().numpy().copy()
        if ref is None:
            ref = out
        same = np.array_equal(ref, out)
        print("warps/CTA cap %2d CTAs/SM %d: %.3f ms  %.3g states/s  world+self %.3f ms  intra-only %.3f ms  %s" %
              (warps, blocks, ms, n / ms * 1e3, run(mb.CHECK_WORLD | mb.CHECK_SELF), run(mb.CHECK_INTRA),
               "" if same else "RESULT DIFFERS"))
eng.set_tuning(0.0, 16, 1)
print("fp64 all states: %.3f ms" % run(mb.CHECK_ALL | mb.CHECK_FP64, 3))
