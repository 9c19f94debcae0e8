#!/usr/bin/env python3
"""Device-resident validity rate of the densified dual-arm robot (PR2 scale: hundreds of spheres, thousands of cluster
pairs).  Usage: rate_dense.py [copies] [n_states]; B200SV_HIER=0/1 selects the flat / hierarchical cull."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402
from test_gpu_parity import _densified  # noqa: E402
from util import RES, ROBOT_FILES, read  # noqa: E402

n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
if len(sys.argv) > 1 and sys.argv[1] in workloads.CONFIGS:  # a named configuration as it is (C2, C3)
    copies, name = 1, sys.argv[1]
    cfg = workloads.CONFIGS[name]
    eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                           resolution=cfg["resolution"], max_distance=cfg["max_distance"])
else:
    copies, name = int(sys.argv[1]) if len(sys.argv) > 1 else 4, "C3"
    u, s = ROBOT_FILES["dualarm"]
    urdf, srdf = read(RES, u), read(RES, s)
    if copies > 1:
        urdf = _densified(urdf, copies)
    cfg = workloads.CONFIGS["C3"]
    eng = mb.StateValidityEngine.from_urdf(urdf, srdf, "arms", cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
eng.field_add_points(workloads.make_cloud(name))
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
import zlib
print("%s x%d hier=%s spheres=%d: %.3f ms per %d states = %.3g states/s, bitmap crc %08x" %
      (name, copies, os.environ.get("B200SV_HIER", "auto"), eng.n_spheres, ms, n, n / ms * 1e3, zlib.crc32(bm.cpu().numpy().tobytes())))
