#!/usr/bin/env python3
"""Device-resident rate of the validity check on a bench workload (default C2), L2 flushed between launches, plus the bitmap's
CRC so variants can be compared: rate_c2.py [workload] [n_states]  (B200SV_LIB selects a library variant)."""
import os
import sys
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = workloads.CONFIGS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else min(cfg["states"]["n"], 2_000_000)
eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                       resolution=cfg["resolution"], max_distance=cfg["max_distance"])
eng.field_add_points(workloads.make_cloud(name))
lo, hi = eng.group_bounds()
q = torch.from_numpy(workloads.random_states(cfg["states"]["seed"], n, lo, hi)).cuda()
bm = torch.zeros(4 * ((n + 31) // 32), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
ts = []
for i in range(13):
    flush.fill_(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    eng.check_batch_device(q.data_ptr(), n, mb.CHECK_ALL, bm.data_ptr(), s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts[3:]))
g = eng.gather_roofline(1 << 28, 3)
print("%s %s: %.4f ms per %d states = %.4g states/s; gather roofline %.0f GB/s -> frac %.3f; rechecked %d; crc %08x" %
      (name, os.environ.get("B200SV_LIB", "default"), ms, n, n / ms * 1e3, g, n / ms * 1e3 * 32 * eng.n_spheres / 1e9 / g,
       eng.stats()["n_rechecked"], zlib.crc32(bm.cpu().numpy().tobytes())))
