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
