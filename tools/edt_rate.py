#!/usr/bin/env python3
"""C4 rebuild (5 M points -> 400 x 400 x 200 voxels) a few times; prints the median device time.  Run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split, or `ncu --set full -k regex:edt` for the kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
c4 = workloads.CONFIGS[name]
e4 = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=c4["size"], origin=c4["origin"],
                                      resolution=c4["resolution"], max_distance=c4["max_distance"])
pts = workloads.make_cloud(name)
d_pts = torch.from_numpy(pts).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
times = []
for i in range(reps):
    e4.field_reset()
    flush.fill_(i)
    ev0.record(stream)
    e4.field_add_points_device(d_pts.data_ptr(), len(pts), mb.FIELD_WORLD, stream.cuda_stream)
    ev1.record(stream)
    torch.cuda.synchronize()
    times.append(ev0.elapsed_time(ev1))
nv = int(np.prod(e4.num_cells))
ms = float(np.median(times[min(3, reps - 1):]))
g4 = e4.field_get_d2()
print("%s: %d points -> %d voxels: %.4f ms = %.3g voxels/s  (min %.4f; d2 checksum %d, %d occupied)" %
      (name, len(pts), nv, ms, nv / ms * 1e3, min(times), int(g4.astype(np.int64).sum()), int((g4 == 0).sum())))
