"""Time of a field build in reference-propagation mode (second build: buffers allocated), with a per-kernel split under ncu.
    python tools/prop_rate.py C2|C4 [reps]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import moveit2_b200 as mb
from moveit2_b200 import workloads

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = workloads.CONFIGS[name]
eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=c["size"], origin=c["origin"], resolution=c["resolution"],
                                       max_distance=c["max_distance"])
pts = workloads.make_cloud(name)
eng.field_set_propagation(mb.PROPAGATION_REFERENCE)
for r in range(reps):
    eng.field_reset()
    t = time.perf_counter()
    eng.field_add_points(pts)
    wall = time.perf_counter() - t
    st = eng.field_propagation_stats()
    print("%s build %d: %.1f ms device, %.1f ms wall; %s" % (name, r, eng.stats()["edt_ms"], wall * 1e3, st), flush=True)
# an object moves: remove 2 % of the points, add them shifted
k = max(1, len(pts) // 50)
t = time.perf_counter()
eng.field_update_points(pts[:k], pts[:k] + 0.03)
print("%s update of %d points: %.1f ms wall; %s" % (name, k, (time.perf_counter() - t) * 1e3, eng.field_propagation_stats()))
eng.close()
