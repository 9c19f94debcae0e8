"""Target for ncu / timing: b200sv_fk_batch_device on 500 000 Panda states (FP32 and FP64 chains, f64 4x4 matrices out)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import moveit2_b200 as mb
from moveit2_b200 import workloads
c = workloads.CONFIGS["C2"]
eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=c["size"], origin=c["origin"], resolution=c["resolution"], max_distance=c["max_distance"])
lo, hi = eng.group_bounds()
n = 500_000
q = torch.from_numpy(workloads.random_states(1, n, lo, hi)).cuda()
T = torch.empty((n, eng.n_links, 16), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
for prec in (32, 64):
    ts = []
    for i in range(6):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        eng.fk_batch_device(q.data_ptr(), n, prec, T.data_ptr(), s.cuda_stream)
        e1.record(s); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    b = 8 * eng.n_group_vars + 128 * eng.n_links
    print("fk fp%d: %.4f ms, %.3g states/s, %.0f GB/s" % (prec, ms, n / ms * 1e3, n * b / ms / 1e6))
eng.close()
