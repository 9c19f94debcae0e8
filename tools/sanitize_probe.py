#!/usr/bin/env python3
"""One small call of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
  compute-sanitizer --tool memcheck python tools/sanitize_probe.py
C1 scene (50^3 field), a few hundred states: EDT (scatter, z+y, x), shape / convex / octree ingest, update + bracket,
the fast FP32 pass + exact recheck, the FP64 pass, the latency path, float32 states, edges, gradients, contacts, FK,
field export / import, live ACM / base-state updates, the gather microbenchmark."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import moveit2_b200 as mb  # noqa: E402
from moveit2_b200 import workloads  # noqa: E402

cfg = workloads.CONFIGS["C1"]
kw = dict(size=cfg["size"], origin=cfg["origin"], resolution=cfg["resolution"], max_distance=cfg["max_distance"])
eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", **kw)
cloud = workloads.make_cloud("C1")
eng.field_add_points(cloud)
eng.field_remove_points(cloud[:100])
eng.field_update_points(cloud[100:300], cloud[200:400] + 0.01)
eng.field_begin_update()
eng.field_add_sphere((0.3, 0.2, 0.4), 0.05)
eng.field_add_shape(mb.SHAPE_BOX, (0.1, 0.08, 0.06), np.eye(4), lattice_frame=mb.LATTICE_BODY)
eng.field_add_shape(mb.SHAPE_CYLINDER, (0.04, 0.1, 0), np.eye(4))
n = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float)
eng.field_add_convex(np.concatenate([n, np.full((6, 1), -0.05)], axis=1), (0, 0, 0, 0.1), np.eye(4))
eng.field_add_octree_leaves(np.array([[0.2, -0.2, 0.5, 0.02], [-0.3, 0.3, 0.3, 0.08]]))
eng.field_end_update()
raw = eng.field_write_df()
eng.field_read_df(raw)
d2 = eng.field_get_d2()
eng.field_set_d2(d2)
lo, hi = eng.group_bounds()
q = workloads.random_states(5, 700, lo, hi)
v = eng.check_batch(q)                                # small-batch path
v64 = eng.check_batch(q, mb.CHECK_ALL | mb.CHECK_FP64)
assert np.array_equal(v, v64)
big = np.concatenate([q] * 8)[:5000 + 7]
assert np.array_equal(eng.check_batch(big)[:700], v)   # chunked large-batch path
assert np.array_equal(eng.check_batch_f32(big.astype(np.float32))[:10], eng.check_batch(big[:10].astype(np.float32).astype(float)))
for i in range(5):
    assert eng.check_batch(q[i:i + 1])[0] == v[i]       # latency path
assert np.array_equal(eng.check_batch(q[:32]), v[:32])
eng.check_motions(q[:100], q[100:200], 0.05 * float(np.sum(hi - lo)))
eng.check_motions(q[:3], q[3:6], 0.05 * float(np.sum(hi - lo)))
eng.collision_gradients(q[:200])
eng.build_link_fields()                                # per-link PosedDistanceFields (tiny grids through the same EDT kernels)
eng.set_gradient_link_mask(np.ones((eng.n_glinks, eng.n_glinks), np.uint8) - np.eye(eng.n_glinks, dtype=np.uint8))
eng.collision_gradients(q[:200])
eng.set_gradient_link_mask(None)
eng.check_contacts(q[:200], max_contacts=4, max_contacts_per_pair=2)
eng.fk_batch(q[:100], 32)
eng.fk_batch(q[:100], 64)
eng.sphere_centers(q[0])
m = eng.model_arrays()
eng.set_acm_masks(m["self_enabled"], m["intra_enabled"])
eng.set_base_state(eng.robot_arrays()["base_positions"])
assert np.array_equal(eng.check_batch(q), v)
eng.gather_roofline(1 << 20, 1)
try:
    import torch
    e2 = mb.StateValidityEngine.for_robot("panda", "panda_arm", **kw)
    buf = torch.empty(eng.field_export_size(), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    eng.field_export_device(buf.data_ptr(), mb.FIELD_WORLD, s)
    e2.field_import_device(buf.data_ptr(), mb.FIELD_WORLD, s)
    torch.cuda.synchronize()
    assert np.array_equal(e2.field_get_d2(), eng.field_get_d2())
    dq = torch.from_numpy(q).cuda()
    bm = torch.zeros(4 * ((len(q) + 31) // 32), dtype=torch.uint8, device="cuda")
    e2.check_batch_device(dq.data_ptr(), len(q), mb.CHECK_ALL, bm.data_ptr(), s)
    torch.cuda.synchronize()
    e2.close()
except ImportError:
    pass
me = mb.MultiDeviceEngine.from_urdf([0, 0], open(os.path.join(mb.engine.RESOURCES, "panda_spheres.urdf")).read(),
                                    open(os.path.join(mb.engine.RESOURCES, "panda.srdf")).read(), "panda_arm",
                                    cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
me.field_add_points(cloud)
me.check_batch(big)
me.close()
# signed field
es = mb.StateValidityEngine.for_robot("panda", "panda_arm", signed=True, **kw)
es.field_add_sphere((0.3, 0.2, 0.4), 0.1)
es.collision_gradients(q[:50])
es.close()
eng.close()
print("sanitize probe ok: %d of %d states valid" % (int(v.sum()), len(v)))
