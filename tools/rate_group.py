#!/usr/bin/env python3
"""Device-resident validity rate of a bundled robot / group on the C3 scene: rate_group.py <robot> <group> [n_states]."""
import os
import sys
import zlib

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

robot, group = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
cfg = workloads.CONFIGS["C3"]
eng = mb.StateValidityEngine.for_robot(robot, group, size=cfg["size"], origin=cfg["origin"], resolution=cfg["resolution"],
                                       max_distance=cfg["max_distance"])
eng.field_add_points(workloads.make_cloud("C3"))
lo, hi = eng.group_bounds()
q = torch.from_numpy(workloads.random_states(41, n, lo, hi)).cuda()
bm = torch.zeros(4 * ((n + 31) // 32), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
for _ in range(3):
    eng.check_batch_device(q.data_ptr(), n, mb.CHECK_ALL, bm.data_ptr(), st.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    eng.check_batch_device(q.data_ptr(), n, mb.CHECK_ALL, bm.data_ptr(), st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("%s %s (%d variables, %d spheres): %.3f ms per %d states = %.3g states/s, bitmap crc %08x, valid %.3f" %
      (robot, group, eng.n_group_vars, eng.n_spheres, ms, n, n / ms * 1e3, zlib.crc32(bm.cpu().numpy().tobytes()),
       float(torch.from_numpy(__import__("numpy").unpackbits(bm.cpu().numpy())).float().mean())))
