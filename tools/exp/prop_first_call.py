import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import moveit2_b200 as mb
from moveit2_b200 import workloads
c = workloads.CONFIGS["C2"]
pts = workloads.make_cloud("C2")
for rep in range(2):
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=c["size"], origin=c["origin"], resolution=c["resolution"], max_distance=c["max_distance"])
    t = time.perf_counter(); eng.field_set_propagation(mb.PROPAGATION_REFERENCE); t1 = time.perf_counter() - t
    t = time.perf_counter(); eng.field_add_points(pts); t2 = time.perf_counter() - t
    eng.field_reset()
    t = time.perf_counter(); eng.field_add_points(pts); t3 = time.perf_counter() - t
    print("resident=%s engine %d: set_propagation %.1f ms, first add %.1f ms (edt_ms %.1f), second add %.1f ms" % (os.environ.get("B200SV_PROP_RESIDENT", "1"), rep, t1 * 1e3, t2 * 1e3, eng.stats()["edt_ms"], t3 * 1e3), flush=True)
    eng.close()
