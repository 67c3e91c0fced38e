"""Parity and size-independent properties at BASELINE.json's full sizes (200 k-triangle tree, 100 k motions, 10^6+
states).  The oracle finishes the 100 k-motion batch in a few seconds on the box's host cores."""
import os

import numpy as np
import pytest

from helpers import BAND, device_robot_from_oracle, parity_report

pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 8


@pytest.fixture(scope="module")
def big(gpu, oracle):
    import bench
    tree = gpu.scenes.make_tree(**bench.TREE)
    mesh = tree.collision_mesh(True)
    model, a, b = bench.make_inputs(gpu, 0)
    return dict(tree=tree, mesh=mesh, dscene=gpu.Scene.from_mesh(mesh), oscene=oracle.Scene(mesh.vertices, mesh.triangles),
                model=model, drb=model.to_device(), orb=oracle.Robot(bench.ARM, bench.JOINTS), a=a, b=b)


def test_config2_100k_motions_parity(gpu, big):
    """BASELINE configs[1] exactly as bench.py runs it: every boolean, step count and toi against the oracle."""
    got = gpu.check_motions(big["dscene"], big["drb"], big["a"], big["b"])
    want = big["oscene"].check_motions(big["orb"], big["a"], big["b"], threads=CORES)
    assert np.array_equal(got["n_steps"], want["n_steps"])
    mism = np.nonzero(got["hit"] != want["hit"])[0]
    if len(mism):  # classify the disagreements: they must all sit inside the 1e-6 m band
        w = big["oscene"].check_motions(big["orb"], big["a"][mism], big["b"][mism], band=True, threads=CORES)
        assert (w["min_abs_clearance"] <= BAND).all(), f"{(w['min_abs_clearance'] > BAND).sum()} mismatches outside the band"
    both = got["hit"] & want["hit"]
    diff_toi = np.nonzero(both & (got["toi"] != want["toi"]))[0]
    if len(diff_toi):  # a differing first colliding step needs a grazing contact at one of the steps up to the later of the two
        w = big["oscene"].check_motions(big["orb"], big["a"][diff_toi], big["b"][diff_toi], band=True, threads=CORES)
        assert (w["min_abs_clearance"] <= BAND).all(), f"{(w['min_abs_clearance'] > BAND).sum()} toi mismatches outside the band"
    assert len(diff_toi) <= 1e-4 * both.sum() + 2
    assert 0.1 < want["hit"].mean() < 0.3


def test_config1_10k_states_parity_on_200k_tree(gpu, oracle, big):
    orb = oracle.Robot(1.0, (0,))
    drb = device_robot_from_oracle(gpu, orb)
    lo, hi = big["tree"].trunk_mesh.aabb()
    h = max(hi[0] - lo[0], hi[1] - lo[1]) / 2 + 2.0
    st = oracle.gen_uniform_states(orb, 10_000, 42, h, lo[2] + 1.0)  # sampling_difficulty.cpp:77-92
    want, cl = big["oscene"].check_states(orb, st, clearance=True, threads=CORES)
    rep = parity_report(gpu.check_states(big["dscene"], drb, st), want, cl)
    assert rep["n_mismatch_outside_band"] == 0, rep


def test_config3_goal_sampling_counts_on_200k_tree(gpu, oracle, big):
    orb = oracle.Robot(1.0, (0,))
    drb = device_robot_from_oracle(gpu, orb)
    targets = big["tree"].fruit_centers[:64]
    got = gpu.sample_goals(big["dscene"], drb, targets, 1000, draws=oracle.goal_draws(orb, 1000, seed=42))
    want = big["oscene"].goal_sample_counts(orb, targets, 1000, seed=42, threads=CORES)
    assert np.abs(got["collisions"].astype(int) - want.astype(int)).max() <= 1  # a count may move by a grazing sample
    assert (got["collisions"] == want).mean() > 0.95


def test_two_million_states_sharding_and_idempotence(gpu, big):
    st = gpu.robots.generate_uniform_random_states(big["model"], gpu.robots.Mt19937(9), 2_000_003, 5.0, 10.0)
    whole = gpu.check_states(big["dscene"], big["drb"], st, packed=True)
    again = gpu.check_states(big["dscene"], big["drb"], st, packed=True)
    assert np.array_equal(whole, again)
    parts = []
    for r in range(8):
        lo, hi = gpu.sharding.shard_range(len(st), r, 8)
        parts.append(gpu.check_states(big["dscene"], big["drb"], st[lo:hi], packed=True))
    assert np.array_equal(gpu.sharding.concat_words(parts, len(st)), whole)
    assert 0.01 < gpu.capi.unpack_bits(whole, len(st)).mean() < 0.1


def test_motion_endpoints_and_first_hit_consistency(gpu, big):
    """Properties of the motion results that need no oracle: a motion whose start state collides has toi == 0; a
    colliding motion's reported step really collides and the step before it does not."""
    a, b = big["a"][:20_000], big["b"][:20_000]
    res = gpu.check_motions(big["dscene"], big["drb"], a, b)
    start_hits = gpu.check_states(big["dscene"], big["drb"], a)
    assert (res["toi"][start_hits] == 0.0).all() and res["hit"][start_hits].all()
    hit = np.nonzero(res["hit"] & ~start_hits)[0][:3000]
    n = res["n_steps"][hit].astype(np.float64)
    k = np.rint(res["toi"][hit] * n)
    assert np.array_equal(k / n, res["toi"][hit])  # toi is exactly step / n_steps

    def interp(t):  # Transform.h:69-78 + RobotState.cpp:19-21 with the Eigen slerp rule, vectorised
        A, B = a[hit], b[hit]
        out = np.empty_like(A)
        out[:, :3] = (1.0 - t)[:, None] * A[:, :3] + t[:, None] * B[:, :3]
        d = (A[:, 3:7] * B[:, 3:7]).sum(1)
        th = np.arccos(np.clip(np.abs(d), 0, 1))
        lin = np.abs(d) >= 1.0 - np.finfo(float).eps
        st = np.where(lin, 1.0, np.sin(th))
        w0 = np.where(lin, 1.0 - t, np.sin((1.0 - t) * th) / st)
        w1 = np.where(lin, t, np.sin(t * th) / st) * np.where(d < 0, -1.0, 1.0)
        out[:, 3:7] = w0[:, None] * A[:, 3:7] + w1[:, None] * B[:, 3:7]
        out[:, 7:] = A[:, 7:] * (1 - t)[:, None] + B[:, 7:] * t[:, None]
        return out
    at_hit = gpu.check_states(big["dscene"], big["drb"], interp(k / n))
    before = gpu.check_states(big["dscene"], big["drb"], interp((k - 1) / n))
    assert at_hit.mean() > 0.999 and before.mean() < 0.001  # numpy's sin differs from the device's by an ulp at most


def _motion_parity(gpu, dscene, drb, oscene, orb, a, b):
    """Booleans, step counts and toi of a motion batch against the oracle, mismatches classified by the clearance band."""
    got = gpu.check_motions(dscene, drb, a, b)
    want = oscene.check_motions(orb, a, b, band=True, threads=CORES)
    assert np.array_equal(got["n_steps"], want["n_steps"])
    in_band = want["min_abs_clearance"] <= BAND
    mism = (got["hit"] != want["hit"]) | (got["hit"] & want["hit"] & (got["toi"] != want["toi"]))
    rep = dict(n_total=len(a), n_mismatch_outside_band=int((mism & ~in_band).sum()), n_in_band=int(in_band.sum()),
               n_mismatch_in_band=int((mism & in_band).sum()), hit_rate=float(want["hit"].mean()))
    assert rep["n_mismatch_outside_band"] == 0, rep
    return rep


def test_config3_goal_sampling_on_the_trunk_scene_robot_sweep(gpu, oracle, big):
    """BASELINE configs[2] on the scene the reference actually collides with (trunk only, fcl_utils.cpp:58-60) for robots of
    the benchmark's sweep (goal_sample_success_rate.cpp:30-38): every sample's boolean against the oracle, band-classified."""
    tmesh = big["tree"].collision_mesh(False)
    dscene, oscene = gpu.Scene.from_mesh(tmesh), oracle.Scene(tmesh.vertices, tmesh.triangles)
    targets, s = big["tree"].fruit_centers[:48], 1000
    valid = {}
    for arm, jt in ((0.5, (0,)), (0.75, (1,)), (1.0, (0, 1)), (1.0, (1, 1))):
        orb = oracle.Robot(arm, jt)
        drb = device_robot_from_oracle(gpu, orb)
        got = gpu.sample_goals(dscene, drb, targets, s, draws=oracle.goal_draws(orb, s, seed=42))
        counts, hits = oscene.goal_sample_counts(orb, targets, s, seed=42, threads=CORES, want_hits=True)
        mism = np.argwhere(got["hit"] != hits)
        for f in np.unique(mism[:, 0]):  # regenerate the disagreeing samples' states and measure their clearance
            st = oracle.gen_goal_states(orb, targets[f], s, seed=42)[mism[mism[:, 0] == f, 1]]
            _, cl = oscene.check_states(orb, st, clearance=True)
            assert (np.abs(cl) <= BAND).all(), (arm, jt, int(f))
        assert len(mism) <= 5 and np.abs(got["collisions"].astype(int) - counts.astype(int)).max() <= 2
        valid[(arm, jt)] = 1.0 - hits.mean()
    # the workload is not degenerate on this scene: a real fraction of the samples is valid, and it depends on the robot
    assert all(0.02 < v < 0.98 for v in valid.values()), valid


def test_config4_prm_edges_full_batch_sampled_parity(gpu, oracle, big):
    """BASELINE configs[3] at its stated size: 100 k roadmap nodes, causal k = 10 neighbours (the reference's incremental
    insertion, tsp_over_prm.cpp:33-66) -> ~1 M edge motions in ONE batch; 30 k of them re-checked by the oracle."""
    cand = gpu.robots.generate_uniform_random_states(big["model"], gpu.robots.Mt19937(7), 104_000, 5.0, 10.0)
    nodes = np.ascontiguousarray(cand[~gpu.check_states(big["dscene"], big["drb"], cand)][:100_000])
    assert len(nodes) == 100_000
    idx, dist = gpu.capi.knn(nodes, 10, causal=True)
    src, col = np.nonzero(idx >= 0)
    assert len(src) == 10 * 100_000 - 55  # node i < 10 has only i predecessors
    a, b = nodes[idx[src, col]], nodes[src]  # motion_collides(neighbour, state), tsp_over_prm.cpp:51-53
    got = gpu.check_motions(big["dscene"], big["drb"], a, b)
    pick = np.random.default_rng(3).choice(len(a), 30_000, replace=False)
    want = big["oscene"].check_motions(big["orb"], a[pick], b[pick], band=True, threads=CORES)
    assert np.array_equal(got["n_steps"][pick], want["n_steps"])
    in_band = want["min_abs_clearance"] <= BAND
    mism = (got["hit"][pick] != want["hit"]) | (got["hit"][pick] & want["hit"] & (got["toi"][pick] != want["toi"]))
    assert not (mism & ~in_band).any(), int((mism & ~in_band).sum())
    # edge weights = equal_weights_distance of the endpoints (tsp_over_prm.cpp:56-58); k-NN spot check against the exhaustive oracle
    some = np.random.default_rng(4).choice(100_000, 64, replace=False)
    for i in some:
        d = np.array([oracle.distance(nodes[i], x) for x in nodes[max(0, i - 4000):i]])  # a window is enough to catch misses cheaply
        if len(d) >= 10:
            assert dist[i, -1] <= np.sort(d)[9] + 1e-12
    d_all = np.array([[oracle.distance(nodes[i], nodes[j]) for j in idx[i]] for i in some[some >= 10]])
    assert np.abs(d_all - dist[some[some >= 10]]).max() < 1e-12
    assert 0.0005 < want["hit"].mean() < 0.2


def test_config5_orchard_parity_sample(gpu, oracle):
    """BASELINE configs[4]: the 8-tree orchard row (~2 M triangles, makeSingleRowOrchard, TreeMeshes.cpp:197-206) with the
    sampler box of SURVEY 8d; 60 k states and 20 k motions against the oracle on the same 2 M triangles."""
    orchard = gpu.scenes.make_orchard()
    mesh = orchard.collision_mesh(True)
    assert 1_900_000 < mesh.n_triangles < 2_200_000
    dscene, oscene = gpu.Scene.from_mesh(mesh), oracle.Scene(mesh.vertices, mesh.triangles)
    orb = oracle.Robot(1.0, (0,))
    drb = device_robot_from_oracle(gpu, orb)
    rng = np.random.default_rng(5)
    n = 100_000
    st = np.zeros((n, 9))
    st[:, 0], st[:, 1], st[:, 2] = rng.uniform(-10, 10, n), rng.uniform(-5, 5, n), rng.uniform(0, 10, n)
    yaw = rng.uniform(0, 2 * np.pi, n)
    st[:, 5], st[:, 6] = np.sin(yaw / 2), np.cos(yaw / 2)
    st[:, 7:] = rng.uniform(-np.pi / 2, np.pi / 2, (n, 2))
    want, cl = oscene.check_states(orb, st[:60_000], clearance=True, threads=CORES)
    rep = parity_report(gpu.check_states(dscene, drb, st[:60_000]), want, cl)
    assert rep["n_mismatch_outside_band"] == 0 and 0.03 < want.mean() < 0.2, rep
    mrep = _motion_parity(gpu, dscene, drb, oscene, orb, st[60_000:80_000], st[80_000:])
    assert 0.15 < mrep["hit_rate"] < 0.6, mrep
    # sharding the same batch (balanced by predicted step counts) changes nothing
    a, b = st[60_000:80_000], st[80_000:]
    whole = gpu.check_motions(dscene, drb, a, b)
    bounds = gpu.sharding.balanced_bounds(gpu.sharding.predicted_motion_costs(a, b), 8)
    parts = [gpu.check_motions(dscene, drb, a[lo:hi], b[lo:hi]) for lo, hi in bounds]
    assert np.array_equal(np.concatenate([p["words"] for p in parts]), whole["words"])
    assert np.array_equal(np.concatenate([p["toi"] for p in parts]), whole["toi"], equal_nan=True)
    costs = gpu.sharding.predicted_motion_costs(a, b)
    assert np.array_equal(costs, whole["n_steps"] + 1)  # the prediction is the step count the device uses
    shard_costs = [costs[lo:hi].sum() for lo, hi in bounds]
    assert max(shard_costs) < 1.02 * np.mean(shard_costs)
