"""GPU parity: the CUDA path (through the C ABI) against the fp64 oracle on identical seeded inputs."""
import numpy as np
import pytest

from helpers import BAND, device_robot_from_oracle, parity_report

pytestmark = pytest.mark.gpu

ROBOTS = [(1.0, (0,)), (0.75, (0,)), (0.5, (1,)), (1.0, (0, 0)), (0.75, (0, 1)), (1.0, (1, 1)), (0.75, (0, 1, 0))]


@pytest.fixture(scope="module")
def tree(mg):
    return mg.scenes.make_tree(seed=7, trunk_triangles=12_000, leaf_triangles=28_000, n_fruit=60)


@pytest.fixture(scope="module", params=[False, True], ids=["T", "T+L"])
def scene_pair(request, gpu, oracle, tree):
    mesh = tree.collision_mesh(request.param)
    return gpu.Scene.from_mesh(mesh), oracle.Scene(mesh.vertices, mesh.triangles), mesh


@pytest.mark.parametrize("arm,jt", ROBOTS)
def test_forward_kinematics_matches_oracle(gpu, oracle, arm, jt):
    orb = oracle.Robot(arm, jt)
    drb = device_robot_from_oracle(gpu, orb)
    states = oracle.gen_uniform_states(orb, 500, seed=3)
    got = gpu.forward_kinematics(drb, states)
    want = np.stack([orb.fk(s) for s in states])
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("arm,jt", ROBOTS[:5])
def test_check_states_matches_oracle(gpu, oracle, scene_pair, arm, jt):
    dscene, oscene, mesh = scene_pair
    orb = oracle.Robot(arm, jt)
    drb = device_robot_from_oracle(gpu, orb)
    lo, hi = mesh.aabb()
    h = max(hi[0] - lo[0], hi[1] - lo[1]) / 2 + 2 * arm  # sampling_difficulty.cpp:77-78
    states = np.concatenate([oracle.gen_uniform_states(orb, 6000, 42, h, lo[2] + arm),
                             oracle.gen_uniform_states(orb, 6000, 43, 2.0, 3.5)])
    want, clearance = oscene.check_states(orb, states, clearance=True, threads=8)
    got = gpu.check_states(dscene, drb, states)
    rep = parity_report(got, want, clearance)
    assert rep["n_mismatch_outside_band"] == 0, rep
    assert 0.02 < want.mean() < 0.98  # the case exercises both outcomes


def test_check_motions_matches_oracle(gpu, oracle, scene_pair):
    dscene, oscene, _ = scene_pair
    orb = oracle.Robot(1.0, (0,))
    drb = device_robot_from_oracle(gpu, orb)
    s = oracle.gen_uniform_states(orb, 3000, 42, 3.0, 4.0)
    a, b = s[:1500], s[1500:]
    want = oscene.check_motions(orb, a, b, band=True, threads=8)
    got = gpu.check_motions(dscene, drb, a, b)
    assert np.array_equal(got["n_steps"], want["n_steps"])
    out_band = want["min_abs_clearance"] > BAND
    assert np.array_equal(got["hit"][out_band], want["hit"][out_band])
    both = got["hit"] & want["hit"] & out_band
    assert np.array_equal(got["toi"][both], want["toi"][both])  # first colliding t, bit for bit
    assert np.isnan(got["toi"][~got["hit"]]).all()
    assert 0.1 < want["hit"].mean() < 0.9


def test_goal_sampling_matches_oracle(gpu, oracle, scene_pair, tree):
    dscene, oscene, _ = scene_pair
    orb = oracle.Robot(1.0, (0,))
    drb = device_robot_from_oracle(gpu, orb)
    targets = tree.fruit_centers[:24]
    draws = oracle.goal_draws(orb, 1000, seed=42)
    got = gpu.sample_goals(dscene, drb, targets, 1000, draws=draws, want_states=True)
    want_counts, want_hits = oscene.goal_sample_counts(orb, targets, 1000, seed=42, threads=8, want_hits=True)
    states = np.stack([oracle.gen_goal_states(orb, t, 1000, seed=42) for t in targets])
    assert np.abs(got["states"] - states).max() < 1e-12
    flat = states.reshape(-1, states.shape[-1])
    _, clearance = oscene.check_states(orb, flat, clearance=True, threads=8)
    rep = parity_report(got["hit"].reshape(-1), want_hits.reshape(-1), clearance)
    assert rep["n_mismatch_outside_band"] == 0, rep
    if rep["n_in_band"] == 0:
        assert np.array_equal(got["collisions"], want_counts)


def test_check_boxes_matches_oracle(gpu, oracle, scene_pair):
    dscene, oscene, _ = scene_pair
    rng = np.random.default_rng(5)
    n = 4000
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    boxes = np.concatenate([rng.uniform([-2, -2, 0], [2, 2, 3.2], (n, 3)), q, rng.uniform(0.02, 0.6, (n, 3))], axis=1)
    want, clearance = oscene.check_obbs(boxes, clearance=True)
    got = gpu.check_boxes(dscene, boxes)
    rep = parity_report(got, want, clearance)
    assert rep["n_mismatch_outside_band"] == 0, rep
    assert 0.05 < want.mean() < 0.95


def _quat(axis, angle):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    return [*(axis * np.sin(angle / 2)), np.cos(angle / 2)]


def build_odd_robot(oracle):
    """A robot the procedural generator never makes: two arms branching off the base, a joint declared child->parent
    (crossed from linkB: the variable transform is inverted, RobotModel.cpp:66-67), attachments with rotations, a link
    with two shapes that carry their own transforms, a fixed joint in the middle of a chain and a link no joint reaches
    (it keeps the root transform, RobotModel.cpp:23-24)."""
    r = oracle.Robot(None)
    base = r.add_link("flying_base")
    r.add_box(base, (0.6, 0.5, 0.2))
    left = r.add_link("left_arm")
    r.add_box(left, (0.05, 0.6, 0.05))
    r.add_box(left, (0.2, 0.05, 0.05), (0.0, 0.3, 0.0, *_quat((0, 0, 1), 0.4)))
    right = r.add_link("right_arm")
    r.add_box(right, (0.04, 0.5, 0.04), (0.0, 0.0, 0.02, 0, 0, 0, 1))
    fore = r.add_link("forearm")
    r.add_box(fore, (0.05, 0.4, 0.05))
    ee = r.add_link("end_effector")
    orphan = r.add_link("orphan")
    r.add_box(orphan, (0.1, 0.1, 0.5), (0.0, 0.0, 0.35, 0, 0, 0, 1))
    r.add_joint(base, left, (0.3, 0.1, 0, *_quat((0, 0, 1), 0.3)), (0, -0.3, 0, 0, 0, 0, 1), oracle.REVOLUTE, (1, 0, 0), -1.2, 1.2)
    # declared from the child's side: link_a = right arm, link_b = base
    r.add_joint(right, base, (0, -0.25, 0, *_quat((1, 0, 0), 0.2)), (-0.3, 0.1, 0, *_quat((0, 1, 0), -0.25)), oracle.REVOLUTE, (0, 0, 1), -1.0, 1.0)
    r.add_joint(left, fore, (0, 0.3, 0, 0, 0, 0, 1), (0, -0.2, 0, *_quat((0, 0, 1), 0.1)), oracle.REVOLUTE, (0, 1, 0), -1.5, 1.5)
    r.add_joint(fore, ee, (0, 0.2, 0, 0, 0, 0, 1), (0, 0, 0, 0, 0, 0, 1), oracle.FIXED)
    return r


def test_general_robot_topologies(gpu, oracle, scene_pair):
    """FK op list, box ordering and reach bound on a branching robot with an inverted joint and an unreachable link."""
    dscene, oscene, _ = scene_pair
    orb = build_odd_robot(oracle)
    drb = device_robot_from_oracle(gpu, orb)
    assert orb.n_variables == drb.n_variables == 3
    rng = np.random.default_rng(8)
    n = 6000
    yaw = rng.uniform(-np.pi, np.pi, n)
    tilt = rng.uniform(-0.4, 0.4, (n, 2))
    q = np.array([oracle.tf_then([0, 0, 0, *_quat((0, 0, 1), y)], [0, 0, 0, *_quat((1, 0, 0), a)])[3:] for y, a in zip(yaw, tilt[:, 0])])
    states = np.concatenate([rng.uniform([-2.2, -2.2, 0.2], [2.2, 2.2, 3.4], (n, 3)), q, rng.uniform(-1.0, 1.0, (n, 3)), rng.uniform(-9, 9, (n, 1))], axis=1)
    got_fk = gpu.forward_kinematics(drb, states[:300])
    want_fk = np.stack([orb.fk(s) for s in states[:300]])
    assert np.abs(got_fk - want_fk).max() < 1e-12
    want, cl = oscene.check_states(orb, states, clearance=True, threads=8)
    rep = parity_report(gpu.check_states(dscene, drb, states), want, cl)
    assert rep["n_mismatch_outside_band"] == 0, rep
    assert 0.05 < want.mean() < 0.95
    a, b = states[:1500], states[1500:3000]
    wm = oscene.check_motions(orb, a, b, band=True, threads=8)
    gm = gpu.check_motions(dscene, drb, a, b)
    assert np.array_equal(gm["n_steps"], wm["n_steps"])  # the surplus 4th joint value counts in the distance
    ok = wm["min_abs_clearance"] > BAND
    assert np.array_equal(gm["hit"][ok], wm["hit"][ok])
    both = gm["hit"] & wm["hit"] & ok
    assert np.array_equal(gm["toi"][both], wm["toi"][both])
    # goal sampling through the forearm chain
    draws = oracle.goal_draws(orb, 400, seed=3)
    assert draws.shape[1] == 4
    targets = np.array([[0.4, 0.3, 1.9], [-0.5, 0.2, 2.3], [3.5, 3.5, 1.0]])
    gg = gpu.sample_goals(dscene, drb, targets, 400, draws=draws, want_states=True)
    want_states = np.stack([oracle.gen_goal_states(orb, t, 400, seed=3) for t in targets])
    assert np.abs(gg["states"] - want_states).max() < 1e-12
    wg, cg = oscene.check_states(orb, want_states.reshape(-1, want_states.shape[-1]), clearance=True, threads=8)
    assert parity_report(gg["hit"].reshape(-1), wg, cg)["n_mismatch_outside_band"] == 0
