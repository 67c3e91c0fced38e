"""Fruit separation: connected components of a mesh's vertex graph (connected_vertex_components,
src/experiment_utils/mesh_connected_components.cpp:25-74).  CPU: the oracle against the reference's own source and a
hand-made case.  GPU: the union-find kernels through the C ABI against the oracle."""
import numpy as np
import pytest


def canonical(labels):
    return np.asarray(labels, dtype=np.uint32)


def test_oracle_components_known_answer(oracle):
    # two triangles sharing vertex 2, one separate triangle, vertex 8 unused
    tris = [[0, 1, 2], [2, 3, 4], [5, 6, 7]]
    assert oracle.mesh_components(tris, 9).tolist() == [0, 0, 0, 0, 0, 5, 5, 5, 8]
    assert oracle.mesh_components(np.zeros((0, 3), np.uint32), 3).tolist() == [0, 1, 2]


def test_oracle_components_match_reference_source(oracle, mg):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    tree = mg.scenes.make_tree(seed=3, trunk_triangles=3000, leaf_triangles=500, n_fruit=5)
    rng = np.random.default_rng(0)
    cases = [(tree.trunk_mesh.triangles, len(tree.trunk_mesh.vertices)), (tree.leaves_mesh.triangles, len(tree.leaves_mesh.vertices)),
             (rng.integers(0, 5000, (3000, 3)), 5000), (rng.integers(0, 300, (2000, 3)), 300)]
    for tris, nv in cases:
        want, n = oracle.ref_mesh_components(tris, nv)
        got = oracle.mesh_components(tris, nv)
        assert np.array_equal(got, want) and len(np.unique(got)) == n


def test_abi_rejects_bad_triangles(mg):
    with pytest.raises(mg.capi.MgodplError, match="out of range"):
        mg.capi.mesh_components([[0, 1, 7]], 3)


@pytest.mark.gpu
def test_gpu_components_equal_oracle(gpu, oracle):
    tree = gpu.scenes.make_tree(seed=9, trunk_triangles=20000, leaf_triangles=30000, n_fruit=10)
    rng = np.random.default_rng(1)
    cases = [(tree.trunk_mesh.triangles, len(tree.trunk_mesh.vertices)), (tree.leaves_mesh.triangles, len(tree.leaves_mesh.vertices)),
             (rng.integers(0, 50000, (30000, 3)), 50000), (rng.integers(0, 100, (5000, 3)), 100), (np.zeros((0, 3), np.uint32), 17),
             (np.array([[3, 3, 3]]), 5)]
    for tris, nv in cases:
        assert np.array_equal(gpu.capi.mesh_components(tris, nv), oracle.mesh_components(tris, nv))
    assert len(gpu.capi.mesh_components(np.zeros((0, 3), np.uint32), 0)) == 0


@pytest.mark.gpu
def test_gpu_components_full_size_properties(gpu):
    """2 M-triangle orchard: labels are idempotent roots, constant across every triangle, and one long strip is one component."""
    orchard = gpu.scenes.make_orchard()
    mesh = orchard.collision_mesh(True) if hasattr(orchard, "collision_mesh") else orchard
    tris, nv = np.asarray(mesh.triangles), len(mesh.vertices)
    labels = gpu.capi.mesh_components(tris, nv)
    assert np.array_equal(labels[labels], labels) and (labels <= np.arange(nv)).all()
    lt = labels[tris]
    assert (lt[:, 0] == lt[:, 1]).all() and (lt[:, 1] == lt[:, 2]).all()
    n = 1_000_000
    strip = np.stack([np.arange(n - 2), np.arange(1, n - 1), np.arange(2, n)], 1)
    assert (gpu.capi.mesh_components(strip, n) == 0).all()
