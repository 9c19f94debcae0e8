"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA path through the C-ABI
(moveit2_b200.StateValidityEngine -> libb200sv.so) and compares with the oracle on the same seeded inputs.
Bar: bit-exact for integer voxel distances and validity booleans; link poses within 1e-5 m (FP32 pass) and
1e-12 m (FP64 pass)."""
import os

import numpy as np
import pytest

import moveit2_b200 as mb
from moveit2_b200 import workloads
from oracle.collision_model import find_internal_points_sphere
from oracle.df import PropagationDistanceField

from util import config_pair, make_pair, unpack

pytestmark = pytest.mark.gpu

W, S, I, ALL, FP64 = mb.CHECK_WORLD, mb.CHECK_SELF, mb.CHECK_INTRA, mb.CHECK_ALL, mb.CHECK_FP64


@pytest.fixture(scope="module")
def c1():
    """Engine whose world/self fields are the REFERENCE's propagation result (injected through
    b200sv_field_set_d2), so validity parity is tested independently of EDT parity."""
    eng, env, cloud, q = config_pair("C1")
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    yield eng, env, cloud, q
    eng.close()


@pytest.fixture(scope="module")
def c1_own():
    """Engine that built its own fields with the CUDA EDT."""
    eng, env, cloud, q = config_pair("C1")
    env.world.add_points(cloud)
    eng.field_add_points(cloud)
    yield eng, env, cloud, q
    eng.close()


# ---- K3 + K4: distance field ----------------------------------------------------------------------------
def test_edt_c1_world_and_self(c1_own):
    """C1's cloud is sparse (2 000 surface points in 125 000 voxels): the reference's vector propagation is an
    approximation there -- it over-estimates some voxels and which ones depends on insertion order -- so
    "bit-exact vs the reference" and "exact" cannot both hold.  The CUDA EDT is exact; every difference must be
    GPU < reference, and they are counted (bench.py reports the count on the benchmark field too)."""
    from scipy import ndimage
    eng, env, cloud, q = c1_own
    assert eng.num_cells == env.world.num_cells and eng.max_distance_sq == env.world.max_distance_sq
    for which, ref in ((mb.FIELD_WORLD, env.world), (mb.FIELD_SELF, env.self_field)):
        g, o = eng.field_get_d2(which), ref.d2()
        assert np.array_equal(g == 0, o == 0)        # identical occupancy (worldToGrid parity, bit-exact)
        exact = np.rint(ndimage.distance_transform_edt(g != 0) ** 2).astype(np.int64)
        assert np.array_equal(g, np.minimum(exact, eng.max_distance_sq))   # the GPU field IS the exact truncated EDT
        assert (g <= o).all()                        # the reference never under-estimates
        frac = float((g != o).mean())
        assert frac < 0.01, frac
        print("field %d: %d of %d voxels differ from the reference propagation (max delta %d)" %
              (which, int((g != o).sum()), g.size, int((o - g).max())))


def exact_edt(occupied, max_d2):
    from scipy import ndimage
    return np.minimum(np.rint(ndimage.distance_transform_edt(~occupied) ** 2).astype(np.int64), max_d2)


def test_edt_dense_cloud_and_reference_order_dependence():
    """Dense clouds (the C4 regime, ~14 % occupancy): the GPU field equals the exact truncated EDT bit for bit.
    The reference's propagation differs from it in a handful of voxels -- and from ITSELF when the same points
    are inserted in another order, which is why no parallel algorithm can be bit-exact against it."""
    eng, env = make_pair("panda", "panda_arm", (0.8, 0.8, 0.4), (0, 0, 0.2), 0.01, 0.25)
    pts = workloads.uniform_cloud(5, 40_000, (0.8, 0.8, 0.4), (0, 0, 0.2))
    eng.field_add_points(pts)
    env.world.add_points(pts)
    g, o = eng.field_get_d2(), env.world.d2()
    assert np.array_equal(g, exact_edt(g == 0, eng.max_distance_sq))
    assert np.array_equal(g == 0, o == 0) and (g <= o).all()
    assert (g != o).sum() <= 16 and (o - g).max() <= 4
    shuffled = PropagationDistanceField(*env.world.args)
    shuffled.add_points(pts[np.random.default_rng(1).permutation(len(pts))])
    order_dependent = int((shuffled.d2() != o).sum())
    print("dense cloud: %d voxels GPU < reference; reference vs its own shuffled-order run: %d voxels" %
          (int((g != o).sum()), order_dependent))
    assert (eng.field_get_d2() == g).all()   # the GPU result is a function of the occupancy SET only
    eng.close()


def test_edt_golden_test_big(golden):
    """test_big.df (reference fixture): sphere r=.5 at (.5,.5,.5) in 150x150x200 @ 2 cm."""
    gz = np.load(os.path.join(golden, "df_golden.npz"))
    dims = tuple(int(v) for v in gz["test_big_dims"])
    occ = np.unpackbits(gz["test_big_occ_bits"])[: dims[0] * dims[1] * dims[2]].reshape(dims).astype(bool)
    eng, env = make_pair("panda", "panda_arm", (3.0, 3.0, 4.0), (1.5, 1.5, 2.0), 0.02, 0.25)
    pts = find_internal_points_sphere((0.5, 0.5, 0.5), 0.5, 0.02)
    eng.field_add_points(pts)
    g = eng.field_get_d2()
    assert np.array_equal(g == 0, occ)
    ref = PropagationDistanceField(3.0, 3.0, 4.0, 0.02, 0, 0, 0, 0.25)
    ref.add_points(pts)
    o = ref.d2()
    assert np.array_equal(g, exact_edt(occ, 169))    # exact
    assert (g <= o).all()                            # reference over-estimates only
    assert (g != o).mean() < 1e-3 and (o - g).max() <= 8
    # the same fixture through the device-side shape ingest (addShapeToField for a sphere): 65 126 voxels, bit for bit
    eng.field_reset()
    eng.field_add_sphere((0.5, 0.5, 0.5), 0.5)
    g2 = eng.field_get_d2()
    assert int((g2 == 0).sum()) == 65126 and np.array_equal(g2 == 0, occ) and np.array_equal(g2, g)
    # a sphere that sticks out of the field: points outside the grid are dropped (isCellValid), as addPointsToField does
    eng.field_reset()
    eng.field_add_sphere((0.05, 1.5, 2.0), 0.3)
    pts2 = find_internal_points_sphere((0.05, 1.5, 2.0), 0.3, 0.02)
    ref2 = PropagationDistanceField(3.0, 3.0, 4.0, 0.02, 0, 0, 0, 0.25)
    ref2.add_points(pts2)
    assert np.array_equal(eng.field_get_d2() == 0, ref2.d2() == 0)
    eng.close()


def test_edt_sparse_scene_mismatch_is_one_sided_and_counted():
    """Sparse scenes: the reference's propagation over-estimates a few voxels (insertion-order dependent,
    SURVEY.md section 7); the GPU field is exact, so every mismatch must be GPU < reference."""
    from scipy import ndimage
    eng, env = make_pair("panda", "panda_arm", (0.8, 0.8, 0.8), (0, 0, 0.4), 0.02, 0.25)
    pts = np.random.default_rng(11).uniform([-0.4, -0.4, 0], [0.4, 0.4, 0.8], size=(60, 3))
    eng.field_add_points(pts)
    env.world.add_points(pts)
    g, o = eng.field_get_d2(), env.world.d2()
    exact = np.minimum(np.rint(ndimage.distance_transform_edt(g != 0) ** 2).astype(np.int64), eng.max_distance_sq)
    assert np.array_equal(g, exact)
    assert (g <= o).all()
    assert (g != o).mean() < 1e-3
    eng.close()


def test_field_add_remove_reset_equals_fresh_build():
    eng, env = make_pair("panda", "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.05, 0.25)
    rng = np.random.default_rng(3)
    a = rng.uniform([-0.5, -0.5, 0], [0.5, 0.5, 1], size=(300, 3))
    b = rng.uniform([-0.5, -0.5, 0], [0.5, 0.5, 1], size=(200, 3))
    eng.field_add_points(a)
    eng.field_add_points(b)          # incremental add == union
    env.world.add_points(np.concatenate([a, b]))
    assert np.array_equal(eng.field_get_d2() == 0, env.world.d2() == 0)
    eng.field_remove_points(a)       # removal == fresh build without those voxels (test_distance_field.cpp:508)
    env.world.remove_points(a)
    fresh, _ = make_pair("panda", "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.05, 0.25)
    occ_b = env.world.d2() == 0
    g = eng.field_get_d2()
    assert np.array_equal(g == 0, occ_b)
    fresh.field_set_d2(np.where(occ_b, 0, eng.max_distance_sq))
    # points outside the grid and empty inputs are ignored
    eng.field_add_points(np.array([[100.0, 0, 0], [0, -100.0, 0]]))
    eng.field_add_points(np.zeros((0, 3)))
    assert np.array_equal(eng.field_get_d2(), g)
    eng.field_reset()
    assert (eng.field_get_d2() == eng.max_distance_sq).all()
    eng.close()
    fresh.close()


def test_field_set_get_roundtrip(c1):
    eng, env = make_pair("panda", "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.02, 0.25)
    o = c1[1].world.d2()
    eng.field_set_d2(o)
    assert np.array_equal(eng.field_get_d2(), o)
    eng.close()


# ---- K1: forward kinematics ------------------------------------------------------------------------------
@pytest.mark.parametrize("robot,group", [("panda", "panda_arm"), ("panda", "panda_arm_hand"), ("dualarm", "arms"),
                                         ("dualarm", "whole_body")])
def test_fk_parity(robot, group):
    eng, env = make_pair(robot, group, (1, 1, 1), (0, 0, 0.5), 0.05, 0.25)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(5, 1000, lo, hi)
    ref = env.fk(q)
    t64 = eng.fk_batch(q, 64)
    t32 = eng.fk_batch(q, 32)
    assert np.abs(t64 - ref).max() < 1e-12       # FP64 pass: same arithmetic, sin/cos within an ulp
    assert np.abs(t32 - ref).max() < 1e-5        # north_star: link poses within 1e-5 m
    assert np.array_equal(t64[:, :, 3, :], np.broadcast_to([0, 0, 0, 1.0], t64[:, :, 3, :].shape))
    # ragged sizes
    for n in (1, 31, 33):
        assert np.abs(eng.fk_batch(q[:n], 64) - ref[:n]).max() < 1e-12
    eng.close()


def test_sphere_centers_parity(c1):
    eng, env, _, q = c1
    for i in range(5):
        assert np.abs(eng.sphere_centers(q[i]) - env.sphere_centers(q[i])).max() < 1e-12


# ---- K2: validity -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [W, S, I, W | S, ALL])
def test_validity_bit_exact_c1(c1, flags):
    eng, env, _, q = c1
    ref, _ = env.check(q, flags)
    for mode in (0, FP64):
        valid = eng.check_batch(q, flags | mode)
        mism = np.nonzero(valid == ref.astype(bool))[0]   # valid must be the negation of collision
        assert len(mism) == 0, "mode %d: %d mismatches, first %s" % (mode, len(mism), mism[:5])
    st = eng.stats()
    assert st["n_states"] == len(q) and st["n_valid"] == int((ref == 0).sum())


def test_validity_with_own_edt_field_differs_one_sidedly(c1_own):
    """With its own (exact) field the GPU can only find MORE collisions than the reference does with its
    over-estimating field; the differing states are counted."""
    eng, env, _, q = c1_own
    ref, _ = env.check(q, ALL)
    valid = eng.check_batch(q, ALL)
    differ = valid == ref.astype(bool)
    assert not (differ & valid).any()       # never "GPU valid, reference colliding"
    assert differ.mean() < 0.01
    print("own-field validity: %d of %d states differ (all GPU-colliding / reference-valid)" %
          (int(differ.sum()), len(q)))


def test_fp32_pass_rechecks_only_a_small_fraction(c1):
    eng, env, _, q = c1
    eng.check_batch(q, ALL)
    st = eng.stats()
    assert st["n_rechecked"] < 0.05 * len(q)


def test_bitmap_layout_and_ragged_batches(c1):
    eng, env, _, q = c1
    ref, _ = env.check(q[:100], ALL)
    for n in (1, 7, 8, 9, 31, 32, 33, 63, 64, 65, 100):
        bm = eng.check_batch_bitmap(q[:n], ALL)
        assert bm.shape == ((n + 7) // 8,)
        assert np.array_equal(unpack(bm, n), ref[:n] == 0)
        if n % 8:
            assert (bm[-1] >> (n % 8)) == 0        # padding bits are zero
    assert eng.check_batch(np.zeros((0, 7)), ALL).shape == (0,)


def test_states_leaving_the_field_are_free():
    """Out-of-bounds spheres read max_distance and never collide (distance_field.cpp:82, types.cpp:229)."""
    eng, env = make_pair("panda", "panda_arm", (0.3, 0.3, 0.3), (2.0, 2.0, 2.0), 0.02, 0.25)
    pts = np.random.default_rng(0).uniform(1.85, 2.15, size=(500, 3))
    env.world.add_points(pts)
    eng.field_set_d2(env.world.d2())
    lo, hi = eng.group_bounds()
    q = workloads.random_states(1, 2000, lo, hi)
    ref, _ = env.check(q, W)
    assert ref.sum() == 0
    assert eng.check_batch(q, W).all()
    eng.close()


def test_tolerance_and_other_resolution():
    eng, env = make_pair("panda", "panda_arm", (1.2, 1.2, 1.2), (0, 0, 0.5), 0.025, 0.3, tolerance=0.01)
    cloud = workloads.clutter_cloud(9, 3000, 8, (-0.5, -0.5, 0.0), (0.5, 0.5, 1.0), 0.22)
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2())
    lo, hi = eng.group_bounds()
    q = workloads.random_states(2, 5000, lo, hi)
    ref, _ = env.check(q, ALL)
    assert 0.05 < ref.mean() < 0.95
    assert np.array_equal(eng.check_batch(q, ALL), ref == 0)
    assert np.array_equal(eng.check_batch(q, ALL | FP64), ref == 0)
    eng.close()


@pytest.mark.parametrize("group", ["arms", "left_arm", "whole_body"])
def test_validity_dualarm(group):
    cfg = workloads.CONFIGS["C3"]
    eng, env = make_pair("dualarm", group, cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    cloud = workloads.make_cloud("C3")
    eng.field_add_points(cloud)
    env.world.add_points(cloud)
    assert np.array_equal(eng.field_get_d2() == 0, env.world.d2() == 0)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)          # validity parity on identical fields
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(4, 20000, lo, hi)
    for flags in (W, S, I, ALL):
        ref, _ = env.check(q, flags)
        assert np.array_equal(eng.check_batch(q, flags), ref == 0), flags
    eng.close()


def test_reference_interface_mirror(c1):
    eng, env, _, q = c1
    cenv = mb.CollisionEnvB200(eng, "panda_arm")
    req = mb.CollisionRequest(group_name="panda_arm")
    ref_all, _ = env.check(q[:50], ALL)
    ref_self, _ = env.check(q[:50], S | I)
    ref_world, _ = env.check(q[:50], W)
    for i in range(50):
        r = mb.CollisionResult()
        cenv.checkCollision(req, r, q[i])
        assert r.collision == bool(ref_all[i])
        r = mb.CollisionResult()
        cenv.checkSelfCollision(req, r, q[i])
        assert r.collision == bool(ref_self[i])
        r = mb.CollisionResult()
        cenv.checkRobotCollision(req, r, q[i])
        assert r.collision == bool(ref_world[i])
    bm = cenv.checkCollisionBatch(req, q[:50])
    assert np.array_equal(unpack(bm, 50), ref_all == 0)


def test_device_resident_entry_point(c1):
    import torch
    eng, env, _, q = c1
    dq = torch.from_numpy(q).cuda()
    bm = torch.zeros(4 * ((len(q) + 31) // 32), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream()
    eng.check_batch_device(dq.data_ptr(), len(q), ALL, bm.data_ptr(), s.cuda_stream)
    s.synchronize()
    ref, _ = env.check(q, ALL)
    assert np.array_equal(unpack(bm.cpu().numpy(), len(q)), ref == 0)
    assert eng.stats()["n_rechecked"] < 0.05 * len(q)


def test_gather_roofline_runs(c1):
    assert c1[0].gather_roofline(1 << 24, 2) > 1.0


# ---- batched edge validation ---------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fraction", [0.01, 0.002])
def test_check_motions_equals_oracle_edge_by_edge(c1, fraction):
    from oracle import motion
    eng, env, cloud, q = c1
    lo, hi = eng.group_bounds()
    L = fraction * float(np.sum(hi - lo))
    rng = np.random.default_rng(7)
    m = 1500
    a = q[:m]
    # a mix of long edges (random targets) and short ones (RRT-style steps), plus a zero-length edge
    b = np.where(rng.random((m, 1)) < 0.5, rng.uniform(lo, hi, size=(m, len(lo))), a + rng.normal(0, 0.05, size=(m, len(lo))))
    b = np.clip(b, lo, hi)
    b[0] = a[0]
    valid, first, n = eng.check_motions(a, b, L)
    rv, rf, rn = motion.check_motions(env, a, b, L)
    assert n == rn
    assert np.array_equal(valid, rv)
    assert np.array_equal(first, rf)
    assert 0 < valid.sum() < m


@pytest.mark.gpu
def test_check_motions_continuous_joints_and_cap(c1):
    from oracle import motion
    eng, env, cloud, q = c1
    lo, hi = eng.group_bounds()
    V = len(lo)
    cont = np.zeros(V, np.uint8)
    cont[0] = 1  # treat joint 1 as continuous: edges across +-pi take the short way round
    fac = np.ones(V)
    fac[3] = 0.0  # and one variable as passive (no contribution to the distance)
    rng = np.random.default_rng(8)
    m = 400
    a = rng.uniform(lo, hi, size=(m, V))
    b = rng.uniform(lo, hi, size=(m, V))
    a[:, 0] = rng.uniform(-np.pi, np.pi, m)
    b[:, 0] = rng.uniform(-np.pi, np.pi, m)
    L = 0.01 * float(np.sum(hi - lo))
    for cap in (0, 5):
        valid, first, n = eng.check_motions(a, b, L, cont, fac, cap)
        rv, rf, rn = motion.check_motions(env, a, b, L, cont, fac, cap)
        assert n == rn and np.array_equal(valid, rv) and np.array_equal(first, rf)


@pytest.mark.gpu
def test_motion_validator_mirror(c1):
    from moveit2_b200.ompl_interface import ModelBasedStateSpace, MotionValidator, StateValidityChecker
    eng, env, cloud, q = c1
    lo, hi = eng.group_bounds()
    space = ModelBasedStateSpace(lo, hi)
    svc, mv = StateValidityChecker(eng, space), MotionValidator(eng, space)
    ok = svc.isValidBatch(q[:256])
    assert svc.isValid(q[0]) == bool(ok[0])
    starts = q[:256][ok]
    valid, frac = mv.checkMotions(starts[:-1], starts[1:])
    assert mv.checkMotion(starts[0], starts[1]) == bool(valid[0])
    assert np.all(frac[valid] == 1.0) and np.all(frac[~valid] < 1.0)
    assert not svc.isValid(hi + 1.0)  # out of bounds is invalid before any collision check


# ---- paths of the fast kernel the default configs do not reach -------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("group", ["hand", "panda_arm_hand"])
def test_validity_prismatic_groups(group):
    """Groups with the (prismatic, mimic) finger joints: the PRISMATIC branch and the parked-centre range check."""
    cfg = workloads.CONFIGS["C1"]
    eng, env = make_pair("panda", group, cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    cloud = workloads.make_cloud("C1")
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(11, 20000, lo, hi)
    q[:50] *= 40.0  # far outside the joint limits: fingers metres away, beyond the travel the host assumed
    for flags in (W, S, I, ALL):
        ref, _ = env.check(q, flags)
        assert np.array_equal(eng.check_batch(q, flags), ref == 0), (group, flags)
    eng.close()


@pytest.mark.gpu
def test_validity_field_covering_part_of_the_workspace():
    """Centres that leave the field on any side read the 1-cell border shell (free); centres just inside do not."""
    eng, env = make_pair("panda", "panda_arm", (0.6, 0.7, 0.5), (0.35, 0.1, 0.45), 0.02, 0.25)
    cloud = workloads.clutter_cloud(21, 4000, 10, (0.08, -0.22, 0.22), (0.62, 0.42, 0.68), 0.0)
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(12, 50000, lo, hi)
    ref, _ = env.check(q, ALL)
    assert 0.02 < ref.mean() < 0.98
    assert np.array_equal(eng.check_batch(q, ALL), ref == 0)
    eng.close()


@pytest.mark.gpu
def test_validity_fine_resolution_uses_16_bit_fields():
    """Thresholds above 255 (sphere radius > 16 cells) switch the compact field to 16 bits per field."""
    eng, env = make_pair("panda", "panda_arm", (0.6, 0.6, 0.6), (0.4, 0.0, 0.5), 0.003, 0.1)
    assert eng.info.field_bytes_per_voxel == 2
    cloud = workloads.clutter_cloud(22, 20000, 6, (0.15, -0.25, 0.25), (0.65, 0.25, 0.75), 0.0)
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(13, 20000, lo, hi)
    ref, _ = env.check(q, ALL)
    assert 0.02 < ref.mean() < 0.98
    assert np.array_equal(eng.check_batch(q, ALL), ref == 0)
    eng.close()


@pytest.mark.gpu
def test_generic_fp32_kernel_agrees_with_fast_kernel(monkeypatch):
    """B200SV_FAST=0 routes the FP32 pass through the generic checkKernel<float> (planar / floating groups use it)."""
    monkeypatch.setenv("B200SV_FAST", "0")
    eng, env, cloud, q = config_pair("C1")
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    ref, _ = env.check(q, ALL)
    assert np.array_equal(eng.check_batch(q, ALL), ref == 0)
    eng.close()


# ---- K5: getCollisionGradients -----------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_collision_gradients_equal_oracle(c1):
    """Per sphere: distance, gradient, CollisionType and per link closest_distance, in the reference's update order."""
    eng, env, cloud, q = c1
    n = 2000
    d, g, t, c, loc = eng.collision_gradients(q[:n])
    rd, rg, rt, rc = env.collision_gradients(q[:n])
    assert d.shape == rd.shape and g.shape == rg.shape
    # sphere centres agree to rounding (the device contracts a*b+c into FMA, the oracle does not), so a centre that sits on a
    # cell face within 1e-15 m may read the neighbouring voxel: such entries are counted, everything else must agree
    both_none = (t == 0) & (rt == 0)
    assert np.array_equal(d[both_none], rd[both_none])          # DBL_MAX
    close = np.isclose(d, rd, rtol=0, atol=1e-9) | both_none
    assert close.mean() > 0.9999, "%d of %d sphere distances differ" % ((~close).sum(), close.size)
    same = close & (t == rt)  # a centre on a cell face (FP64 rounding) can change which source wins a tie
    assert same.mean() > 0.9995
    assert np.allclose(g[same], rg[same], rtol=0, atol=1e-9)
    assert np.isclose(c, rc, rtol=0, atol=1e-9).mean() > 0.999
    for k in (1, 2, 3):
        assert (rt == k).any(), "the test scene should exercise CollisionType %d" % k
    # sphere_locations are the posed centres
    assert np.allclose(loc[0], env.sphere_centers(q[0]), atol=1e-12)
    # intra gradients are un-normalised centre differences: their norm is the recorded distance
    intra = t == 2
    assert np.allclose(np.linalg.norm(g[intra], axis=1), d[intra], atol=1e-12)


# ---- attached bodies -----------------------------------------------------------------------------------------------------
def _with_attached_bodies():
    cfg = workloads.CONFIGS["C1"]
    from oracle.collision_model import OracleEnv
    from oracle.robot_model import RobotModel
    from util import RES, ROBOT_FILES, read
    u, s = ROBOT_FILES["panda"]
    env = OracleEnv(RobotModel(read(RES, u), read(RES, s)), "panda_arm", cfg["size"], cfg["origin"], cfg["resolution"],
                    cfg["max_distance"])
    # a long rod held in the hand (two shapes = two entries, mutually disabled by the same name) and a box on the forearm
    rod = [((0.0, 0.0, 0.20 + 0.05 * k), 0.025) for k in range(6)]  # clear of the fingers (not touch links of the hand)
    bodies = [dict(link="panda_hand", name="rod", spheres=rod[:3], bs_center=(0, 0, 0.25), bs_radius=0.08,
                   touch_links={"panda_hand", "panda_leftfinger", "panda_rightfinger"}),
              dict(link="panda_hand", name="rod", spheres=rod[3:], bs_center=(0, 0, 0.40), bs_radius=0.08,
                   touch_links={"panda_hand"}),
              dict(link="panda_link5", name="box", spheres=[((0.0, 0.16, -0.1), 0.04), ((0.0, 0.16, -0.02), 0.04)],
                   bs_center=(0.0, 0.16, -0.06), bs_radius=0.09, touch_links={"panda_link5"})]
    env.attach_bodies(bodies)
    # same name => same body: the reference never tests a body against itself
    n0 = len(env.group.updated_links)
    intra = env.arr["intra_enabled"].reshape(n0 + 3, n0 + 3)
    intra[n0, n0 + 1] = 0
    env.arr["intra_enabled"][:] = intra.reshape(-1)
    robot, model = env.flat_arrays()
    eng = mb.StateValidityEngine.from_arrays(robot, model, cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    return eng, env


@pytest.mark.gpu
def test_validity_with_attached_bodies():
    eng, env = _with_attached_bodies()
    assert eng.n_spheres == env.n_spheres == 58 + 8
    cloud = workloads.make_cloud("C1")
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = env.joint_bounds()
    q = workloads.random_states(31, 30000, lo, hi)
    base, _ = env.check(q, I)
    assert 0.01 < base.mean() < 0.9
    for flags in (W, S, I, ALL):
        ref, _ = env.check(q, flags)
        for mode in (0, FP64):
            assert np.array_equal(eng.check_batch(q, flags | mode), ref == 0), (flags, mode)
    # the bodies matter: more intra-group collisions than the bare arm has
    bare_eng, bare_env, _, _ = config_pair("C1", n_states=1)
    bare, _ = bare_env.check(q, I)
    assert base.sum() > bare.sum()
    d, g, t, c, loc = eng.collision_gradients(q[:500])
    rd, rg, rt, rc = env.collision_gradients(q[:500])
    close = np.isclose(d, rd, rtol=0, atol=1e-9) | ((t == 0) & (rt == 0))
    # (a centre on a cell face to within FP64 rounding can change WHICH source wins a tie without changing the distance)
    assert close.mean() > 0.9999 and (t[close] == rt[close]).mean() > 0.9995
    eng.close()
    bare_eng.close()


# ---- .df streams (writeToStream / readFromStream) -----------------------------------------------------------------------
@pytest.mark.gpu
def test_df_stream_round_trip_and_reference_format():
    from oracle.df import read_df_occupancy, write_df_occupancy
    eng, env, cloud, q = config_pair("C1", n_states=1)
    eng.field_add_points(cloud)
    before = eng.field_get_d2()
    raw = eng.field_write_df()
    hdr, n, occ = read_df_occupancy(raw)  # the oracle's reader of the reference format
    assert tuple(n) == eng.num_cells and abs(hdr["resolution"] - 0.02) < 1e-12
    assert abs(hdr["origin_z"] - 0.0) < 1e-12 and abs(hdr["origin_x"] + 0.5) < 1e-12  # the grid CORNER
    mask = np.zeros(eng.num_cells, bool)
    mask[occ[:, 0], occ[:, 1], occ[:, 2]] = True
    assert np.array_equal(mask.reshape(-1), (before == 0).reshape(-1))
    eng.field_reset()
    assert (eng.field_get_d2() != 0).all()
    eng.field_read_df(raw)
    assert np.array_equal(eng.field_get_d2(), before)
    # a stream produced by the Python writer of the same format (what the reference's fixtures look like)
    occ2 = np.zeros(eng.num_cells, bool)
    occ2[3, 4, 5] = occ2[10, 20, 49] = occ2[49, 0, 7] = True
    eng.field_read_df(write_df_occupancy(hdr, occ2))
    d2 = eng.field_get_d2().reshape(eng.num_cells)
    assert np.array_equal(np.argwhere(d2 == 0), np.argwhere(occ2))
    # geometry mismatch is an error, not a silent resize
    bad = dict(hdr)
    bad["resolution"] = 0.05
    with pytest.raises(mb.B200svError):
        eng.field_read_df(write_df_occupancy(bad, occ2))
    eng.close()


# ---- contacts (checkCollision with req.contacts) ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("limits", [(1, 1), (4, 2), (10, 3), (64, 64)])
def test_contacts_equal_oracle(c1, limits):
    """Which contacts are reported (and how many) under max_contacts / max_contacts_per_pair is the reference's choice."""
    eng, env, cloud, q = c1
    mc, mp = limits
    n = 4000
    for flags in (ALL, S | I, W):
        col, cnt, con = eng.check_contacts(q[:n], flags, mc, mp)
        rcol, rcnt, rcon = env.check_contacts(q[:n], flags, mc, mp)
        assert np.array_equal(col, rcol) and np.array_equal(cnt, rcnt), flags
        ref, _ = env.check(q[:n], flags)
        assert np.array_equal(col, ref)  # res.collision is the same boolean with or without contacts
        for i in np.nonzero(cnt)[0]:
            a, b = con[i, : cnt[i]], rcon[i, : cnt[i]]
            for f in ("body_1", "body_2", "sphere_1", "sphere_2"):
                assert np.array_equal(a[f], b[f]), (flags, i, f)
            assert np.allclose(a["pos"], b["pos"], atol=1e-12)
        assert cnt.max() <= mc and (cnt[col == 0] == 0).all()


@pytest.mark.gpu
def test_contact_map_of_the_mirror(c1):
    eng, env, cloud, q = c1
    cenv = mb.CollisionEnvB200(eng, "panda_arm")
    req = mb.CollisionRequest(group_name="panda_arm", contacts=True, max_contacts=8, max_contacts_per_pair=2)
    ref, _ = env.check(q[:200], ALL)
    seen = set()
    for i in np.nonzero(ref[:200])[0][:30]:
        r = mb.CollisionResult()
        cenv.checkCollision(req, r, q[i])
        assert r.collision and 0 < r.contact_count <= 8
        assert sum(len(v) for v in r.contacts.values()) == r.contact_count
        assert all(len(v) <= 2 for v in r.contacts.values())
        seen |= {k[1] for k in r.contacts}
    assert "environment" in seen


@pytest.mark.gpu
def test_empty_and_invalid_inputs_of_the_batched_entry_points(c1):
    eng, env, cloud, q = c1
    V = eng.n_group_vars
    valid, first, n = eng.check_motions(np.zeros((0, V)), np.zeros((0, V)), 0.1)
    assert valid.shape == (0,) and n == 0
    d, g, t, c, loc = eng.collision_gradients(np.zeros((0, V)))
    assert d.shape == (0, eng.n_spheres)
    col, cnt, con = eng.check_contacts(np.zeros((0, V)), max_contacts=3)
    assert col.shape == (0,) and con.shape == (0, 3)
    with pytest.raises(mb.B200svError):
        eng.check_motions(q[:1], q[1:2], 0.0)       # longest valid segment must be positive
    with pytest.raises(mb.B200svError):
        eng.check_contacts(q[:1], max_contacts=0)
    with pytest.raises(mb.B200svError):
        eng.check_contacts(q[:1], flags=0, max_contacts=1)
    # one edge whose end points coincide: only s2 is checked
    valid, first, n = eng.check_motions(q[:1], q[:1], 0.1)
    assert n == 1 and bool(valid[0]) == bool(eng.check_batch(q[:1])[0])


@pytest.mark.gpu
def test_two_contexts_and_threads(c1):
    """CollisionEnv check methods are const and called concurrently (planning_scene/test/test_multi_threaded.cpp:98-140):
    two contexts (different robots) driven from four host threads give the same answers as serial calls."""
    import threading
    eng, env, cloud, q = c1
    cfg = workloads.CONFIGS["C3"]
    eng2, env2 = make_pair("dualarm", "arms", cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    cloud2 = workloads.make_cloud("C3")
    eng2.field_add_points(cloud2)
    lo, hi = eng2.group_bounds()
    q2 = workloads.random_states(5, 20000, lo, hi)
    want1, want2 = eng.check_batch(q), eng2.check_batch(q2)
    out, errs = {}, []

    def work(name, e, qq, reps):
        try:
            for r in range(reps):
                out[(name, r)] = e.check_batch(qq)
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    threads = [threading.Thread(target=work, args=("a%d" % i, eng, q, 5)) for i in range(2)] + \
              [threading.Thread(target=work, args=("b%d" % i, eng2, q2, 5)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs
    for (name, r), v in out.items():
        assert np.array_equal(v, want1 if name.startswith("a") else want2), (name, r)
    eng2.close()


@pytest.mark.gpu
def test_field_update_points_equals_reference_set_semantics():
    """updatePointsInField (propagation_distance_field.cpp:130-175): occupancy afterwards = the oracle's, incl. the
    corner case of a voxel that is in both sets but not currently occupied (left alone)."""
    eng, env, cloud, q = config_pair("C1", n_states=1)
    rng = np.random.default_rng(3)
    first = cloud[:1500]
    eng.field_add_points(first)
    env.world.add_points(first)
    old = np.concatenate([first[:600], rng.uniform(-0.4, 0.4, size=(50, 3)) + (0, 0, 0.5)])   # 50 never added
    new = np.concatenate([first[400:900] + rng.normal(0, 0.03, size=(500, 3)), old[-50:-25]])  # 25 in both, unoccupied
    eng.field_update_points(old, new)
    env.world.update_points(old, new)
    g, o = eng.field_get_d2(), env.world.d2()
    assert np.array_equal(g == 0, o == 0)
    # Distances: the GPU field is the exact transform of that occupancy.  The reference's incremental removal
    # (removeObstacleVoxels :303-388, restated line by line in the oracle) re-propagates only through voxels it could
    # invalidate by flooding from the removed ones, and leaves stale UNDER-estimates behind (counted here); a fresh
    # reference build of the same occupancy never under-estimates.
    from scipy import ndimage
    exact = np.minimum(np.rint(ndimage.distance_transform_edt(g != 0) ** 2).astype(np.int64), eng.max_distance_sq)
    assert np.array_equal(g, exact)
    print("updatePointsInField: %d of %d reference voxels are stale under-estimates" % (int((o < g).sum()), g.size))
    eng.close()


# ---- a robot at PR2 scale: many spheres per link, thousands of cluster pairs ---------------------------------------------
def _densified(urdf, copies):
    """Every collision sphere of the URDF replaced by `copies` slightly smaller ones spread around it."""
    import re
    offs = [(0.0, 0.0, 0.0), (0.015, 0.0, 0.0), (-0.015, 0.0, 0.0), (0.0, 0.015, 0.0), (0.0, -0.015, 0.0), (0.0, 0.0, 0.015)]

    def rep(m):
        x, y, z = (float(v) for v in m.group(1).split())
        r = float(m.group(2))
        return "".join('<collision><origin rpy="0 0 0" xyz="%r %r %r"/><geometry><sphere radius="%r"/></geometry></collision>'
                       % (x + o[0], y + o[1], z + o[2], r - 0.01) for o in offs[:copies])

    return re.sub(r'<collision><origin rpy="0 0 0" xyz="([^"]+)"/><geometry><sphere radius="([^"]+)"/></geometry></collision>',
                  rep, urdf)


@pytest.mark.gpu
def test_validity_dense_dual_arm():
    from oracle.collision_model import OracleEnv
    from oracle.robot_model import RobotModel
    from util import RES, ROBOT_FILES, read
    u, s = ROBOT_FILES["dualarm"]
    urdf, srdf = _densified(read(RES, u), 4), read(RES, s)
    cfg = workloads.CONFIGS["C3"]
    eng = mb.StateValidityEngine.from_urdf(urdf, srdf, "arms", cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    env = OracleEnv(RobotModel(urdf, srdf), "arms", cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    assert eng.n_spheres == env.n_spheres >= 250
    cloud = workloads.make_cloud("C3")
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(41, 30000, lo, hi)
    for flags in (W, S | I, ALL):
        ref, _ = env.check(q, flags)
        assert np.array_equal(eng.check_batch(q, flags), ref == 0), flags
    assert 0.05 < ref.mean() < 0.95
    eng.close()


# ---- signed fields (use_signed_distance_field_ / propagate_negative_) ------------------------------------------------------
def _solid_scene():
    """Solid bodies the arm can reach into: spheres voxelised with the reference's own lattice (findInternalPointsConvex)."""
    return [((0.45, 0.15, 0.55), 0.17), ((0.3, -0.35, 0.35), 0.14), ((-0.1, 0.4, 0.7), 0.12), ((0.42, 0.2, 0.5), 0.1)]


@pytest.mark.gpu
def test_signed_field_edt_is_the_exact_transform_of_the_complement():
    from scipy import ndimage
    cfg = workloads.CONFIGS["C1"]
    eng, env = make_pair(cfg["robot"], cfg["group"], cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"], signed=True)
    for c, r in _solid_scene():
        eng.field_add_sphere(c, r)
        env.world.add_points(find_internal_points_sphere(c, r, cfg["resolution"]))
    g, gn, o, on = eng.field_get_d2(), eng.field_get_neg_d2(), env.world.d2(), env.world.neg_d2()
    occ = g == 0
    assert np.array_equal(occ, o == 0) and occ.sum() > 3000
    exact = np.minimum(np.rint(ndimage.distance_transform_edt(occ) ** 2).astype(np.int64), eng.max_distance_sq)
    assert np.array_equal(gn, exact)          # 0 in free voxels, distance to the nearest free voxel inside obstacles
    assert gn.max() >= 16                     # the scene is deep enough to mean something
    # the reference's propagateNegative never under-estimates and differs in few voxels (counted)
    assert (on >= gn).all() and np.array_equal(on > 0, occ)
    differing = int((on != gn).sum())
    assert differing <= 0.05 * occ.sum()
    print("signed field: %d occupied voxels, %d of them above the exact value in the reference's propagation" % (occ.sum(), differing))
    # the self field (voxelised non-group links) gets its negative half too
    sn = eng.field_get_neg_d2(mb.FIELD_SELF)
    assert np.array_equal(sn > 0, eng.field_get_d2(mb.FIELD_SELF) == 0) and (env.self_field.neg_d2() >= sn).all()
    # removal == fresh build (test_distance_field.cpp:508), reset clears both halves
    c0, r0 = _solid_scene()[0]
    eng.field_remove_points(find_internal_points_sphere(c0, r0, cfg["resolution"]))
    fresh, _ = make_pair(cfg["robot"], cfg["group"], cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"], signed=True)
    for c, r in _solid_scene()[1:]:
        fresh.field_add_sphere(c, r)
    # (sphere 3 overlaps sphere 0: the voxels they share are removed with sphere 0, as in the reference)
    fresh.field_remove_points(find_internal_points_sphere(c0, r0, cfg["resolution"]))
    assert np.array_equal(eng.field_get_d2(), fresh.field_get_d2()) and np.array_equal(eng.field_get_neg_d2(), fresh.field_get_neg_d2())
    eng.field_reset()
    assert not eng.field_get_neg_d2().any() and (eng.field_get_d2() == eng.max_distance_sq).all()
    eng.close()
    fresh.close()


@pytest.mark.gpu
def test_signed_field_validity_and_gradients_equal_oracle():
    """Signed fields change distances and gradients inside obstacles (CHOMP's way out of a collision), not the boolean."""
    cfg = workloads.CONFIGS["C1"]
    eng, env = make_pair(cfg["robot"], cfg["group"], cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"], signed=True)
    uns, _ = make_pair(cfg["robot"], cfg["group"], cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    for c, r in _solid_scene():
        env.world.add_points(find_internal_points_sphere(c, r, cfg["resolution"]))
    # the reference's own propagation result, both halves, so this test is independent of EDT parity
    for which, f in ((mb.FIELD_WORLD, env.world), (mb.FIELD_SELF, env.self_field)):
        eng.field_set_d2(f.d2(), which)
        eng.field_set_neg_d2(f.neg_d2(), which)
        uns.field_set_d2(f.d2(), which)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(77, 20000, lo, hi)
    for flags in (W, S, ALL):
        ref, _ = env.check(q, flags)
        assert np.array_equal(eng.check_batch(q, flags), ref == 0), flags
        assert np.array_equal(uns.check_batch(q, flags), ref == 0), flags
    assert 0.05 < ref.mean() < 0.95
    n = 3000
    d, g, t, c, loc = eng.collision_gradients(q[:n])
    rd, rg, rt, rc = env.collision_gradients(q[:n])
    both_none = (t == 0) & (rt == 0)
    close = np.isclose(d, rd, rtol=0, atol=1e-9) | both_none
    assert close.mean() > 0.9999, "%d of %d sphere distances differ" % ((~close).sum(), close.size)
    same = close & (t == rt)
    assert same.mean() > 0.9995
    assert np.allclose(g[same], rg[same], rtol=0, atol=1e-9)
    assert np.isclose(c, rc, rtol=0, atol=1e-9).mean() > 0.999
    # spheres inside an obstacle carry negative distances, which the unsigned field cannot express
    inside = (rt == 3) & (rd < 0)
    assert inside.sum() > 100 and (d[inside & same] < 0).all()
    du = uns.collision_gradients(q[:n])[0]
    assert (du[inside] >= 0).all()
    eng.close()
    uns.close()


# ---- world ingest: spheres, cylinders and boxes decomposed on the device ---------------------------------------------------
def _pose(rpy, xyz):
    cr, sr, cp, sp, cy, sy = np.cos(rpy[0]), np.sin(rpy[0]), np.cos(rpy[1]), np.sin(rpy[1]), np.cos(rpy[2]), np.sin(rpy[2])
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, xyz
    return T


@pytest.mark.gpu
@pytest.mark.parametrize("frame", ["world", "body"])
def test_field_add_shape_equals_the_restated_body_decomposition(frame):
    from oracle.collision_model import find_internal_points_body
    cfg = workloads.CONFIGS["C1"]
    eng, env = make_pair(cfg["robot"], cfg["group"], cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    lf = mb.LATTICE_WORLD if frame == "world" else mb.LATTICE_BODY
    shapes = [(mb.SHAPE_BOX, (0.21, 0.13, 0.37), _pose((0.3, -0.7, 1.1), (0.31, 0.12, 0.48))),
              (mb.SHAPE_CYLINDER, (0.09, 0.43), _pose((1.2, 0.4, -0.5), (-0.2, 0.25, 0.6))),
              (mb.SHAPE_SPHERE, (0.11,), _pose((0, 0, 0), (0.05, -0.33, 0.3))),
              (mb.SHAPE_BOX, (0.3, 0.3, 0.06), _pose((0, 0, 0.4), (0.1, 0.1, 0.97)))]   # sticks out of the field: clipped
    total = 0
    for ty, dims, T in shapes:
        pts = find_internal_points_body(ty, dims, T, cfg["resolution"], frame)
        assert len(pts) > 200
        total += len(pts)
        env.world.add_points(pts)
        eng.field_add_shape(ty, dims, T, lattice_frame=lf)
        assert np.array_equal(eng.field_get_d2() == 0, env.world.d2() == 0), (ty, frame)
    g, o = eng.field_get_d2(), env.world.d2()
    assert (g <= o).all() and (g == 0).sum() > 0.5 * total / (1 if frame == "world" else 2)
    # removeShapeFromField: clears exactly the voxels of the body's points
    ty, dims, T = shapes[0]
    env.world.remove_points(find_internal_points_body(ty, dims, T, cfg["resolution"], frame))
    eng.field_add_shape(ty, dims, T, lattice_frame=lf, remove=True)
    assert np.array_equal(eng.field_get_d2() == 0, env.world.d2() == 0)
    # the sphere in the world frame is the construction the reference's test_big.df pins
    if frame == "world":
        ref = find_internal_points_sphere((0.05, -0.33, 0.3), 0.11, cfg["resolution"])
        assert np.array_equal(ref, find_internal_points_body(1, (0.11,), shapes[2][2], cfg["resolution"], "world"))
    with pytest.raises(mb.B200svError):
        eng.field_add_shape(3, (0.1, 0.2), np.eye(4))  # cone: host side
    eng.close()


# ---- link geometry given as cylinders and boxes (BodyDecomposition's bounding-cylinder spheres) ----------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("group", ["panda_arm", "panda_arm_hand"])
def test_validity_with_cylinder_and_box_link_geometry(group):
    cfg = workloads.CONFIGS["C1"]
    eng, env = make_pair("panda_primitives", group, cfg["size"], cfg["origin"], cfg["resolution"], cfg["max_distance"])
    assert eng.n_spheres == env.n_spheres == 52
    cloud = workloads.make_cloud("C1")
    eng.field_add_points(cloud)
    env.world.add_points(cloud)
    # the self field holds the voxelised BOX of panda_link0 (interior lattice of a non-sphere body)
    assert np.array_equal(eng.field_get_d2(mb.FIELD_SELF) == 0, env.self_field.d2() == 0)
    assert (env.self_field.d2() == 0).sum() > 300
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(9, 20000, lo, hi)
    for flags in (W, S, I, ALL):
        ref, _ = env.check(q, flags)
        assert np.array_equal(eng.check_batch(q, flags), ref == 0), flags
    assert 0.05 < ref.mean() < 0.95
    for s in q[:8]:
        assert np.abs(eng.sphere_centers(s) - env.sphere_centers(s).reshape(-1, 3)).max() < 1e-12
    eng.close()


# ---- a floating joint in the group (a free-flying base): floating_joint_model.cpp:231-236 ------------------------------------
@pytest.mark.gpu
def test_validity_and_fk_with_a_floating_joint_in_the_group():
    from oracle.collision_model import OracleEnv
    from oracle.robot_model import RobotModel
    from util import RES, ROBOT_FILES, read
    u, s = ROBOT_FILES["panda"]
    urdf = read(RES, u)
    srdf = read(RES, s).replace("</robot>", '<group name="panda_free"><joint name="virtual_joint"/><group name="panda_arm"/></group></robot>')
    cfg = workloads.CONFIGS["C2"]
    eng = mb.StateValidityEngine.from_urdf(urdf, srdf, "panda_free", cfg["size"], cfg["origin"], 0.02, cfg["max_distance"])
    env = OracleEnv(RobotModel(urdf, srdf), "panda_free", cfg["size"], cfg["origin"], 0.02, cfg["max_distance"])
    assert eng.n_group_vars == 14
    cloud = workloads.make_cloud("C2")[::10]
    env.world.add_points(cloud)
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    q = workloads.random_states(12, 30000, lo, hi)   # translations in a 1 m box, un-normalised quaternions (the joint normalises)
    t = eng.fk_batch(q[:64], 64)
    assert np.abs(t - env.fk(q[:64])).max() < 1e-9
    assert np.abs(eng.fk_batch(q[:64], 32) - env.fk(q[:64])).max() < 1e-5
    for flags in (W, S, I, ALL):
        ref, _ = env.check(q, flags)
        assert np.array_equal(eng.check_batch(q, flags), ref == 0), flags
    assert 0.05 < ref.mean() < 0.95
    eng.close()
