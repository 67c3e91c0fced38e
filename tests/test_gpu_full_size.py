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
    same_toi = got["toi"][both] == want["toi"][both]
    assert same_toi.mean() > 0.9999  # a differing toi needs a grazing contact at an earlier step
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
