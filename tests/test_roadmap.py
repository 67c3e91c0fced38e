"""Roadmap shortest paths (runDijkstra per goal sample, tsp_over_prm.cpp:80-97, 497-517).

CPU part: the oracle's Dijkstra against hand-made answers and against scipy's implementation (an independent third
party; Boost.Graph itself is absent, so this row's parity is otherwise unpinned).  GPU part: the batched
label-correcting kernels through the C ABI, bit-equal distances, valid predecessors."""
import numpy as np
import pytest

DBL_MAX = np.finfo(np.float64).max


def random_graph(rng, n, e, lo=0.05, hi=3.0):
    edges = rng.integers(0, n, (e, 2)).astype(np.uint32)
    return edges, rng.uniform(lo, hi, e)


def scipy_distances(n, edges, weights, sources):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import dijkstra
    best = {}
    for (a, b), w in zip(edges.tolist(), weights.tolist()):
        if a != b:
            k = (min(a, b), max(a, b))
            best[k] = min(best.get(k, np.inf), w)
    r = np.array([k[0] for k in best], dtype=np.int64)
    c = np.array([k[1] for k in best], dtype=np.int64)
    v = np.array(list(best.values()))
    m = coo_matrix((np.r_[v, v], (np.r_[r, c], np.r_[c, r])), shape=(n, n)).tocsr()
    d = dijkstra(m, directed=True, indices=np.asarray(sources, dtype=np.int64))
    d[np.isinf(d)] = DBL_MAX
    return d


def check_predecessors(n, edges, weights, sources, dist, pred):
    """pred[v] is a neighbour with dist[pred] + w == dist[v]; sources and unreachable vertices point at themselves."""
    adj = {}
    for (a, b), w in zip(edges.tolist(), weights.tolist()):
        adj.setdefault((a, b), []).append(w)
        adj.setdefault((b, a), []).append(w)
    for s, src in enumerate(np.asarray(sources).tolist()):
        for v in range(n):
            p = int(pred[s, v])
            if v == src or dist[s, v] == DBL_MAX:
                assert p == v
            else:
                assert p != v and any(dist[s, p] + w == dist[s, v] for w in adj[(p, v)]), (s, v, p)


# ---- oracle (CPU) ----------------------------------------------------------------------------------------------------
def test_oracle_dijkstra_known_answers(oracle):
    #   0 --1.0-- 1 --2.0-- 2      4 isolated; edge 0-2 costs 3.5 (longer than 0-1-2), 2-3 costs 0.25
    edges = np.array([[0, 1], [1, 2], [0, 2], [2, 3]], np.uint32)
    w = np.array([1.0, 2.0, 3.5, 0.25])
    d, p = oracle.roadmap_shortest_paths(5, edges, w, [0, 3])
    assert d[0].tolist() == [0.0, 1.0, 3.0, 3.25, DBL_MAX] and p[0].tolist() == [0, 0, 1, 2, 4]
    assert d[1].tolist() == [3.25, 2.25, 0.25, 0.0, DBL_MAX] and p[1].tolist() == [1, 2, 3, 3, 4]


def test_oracle_dijkstra_matches_scipy(oracle):
    rng = np.random.default_rng(0)
    for n, e in ((50, 60), (400, 1500), (3000, 12000)):
        edges, w = random_graph(rng, n, e)
        src = rng.integers(0, n, 24).astype(np.uint32)
        d, p = oracle.roadmap_shortest_paths(n, edges, w, src, threads=4)
        assert np.array_equal(d, scipy_distances(n, edges, w, src))
        check_predecessors(n, edges, w, src, d, p)


def test_abi_validates_graph_arguments(mg):
    """Argument errors are reported before any device work, so this runs without a GPU."""
    with pytest.raises(mg.capi.MgodplError, match="non-negative"):
        mg.capi.roadmap_shortest_paths(3, [[0, 1]], [-1.0], [0])
    with pytest.raises(mg.capi.MgodplError, match="non-negative"):
        mg.capi.roadmap_shortest_paths(3, [[0, 1]], [np.nan], [0])
    with pytest.raises(mg.capi.MgodplError, match="out of range"):
        mg.capi.roadmap_shortest_paths(3, [[0, 3]], [1.0], [0])
    with pytest.raises(mg.capi.MgodplError, match="out of range"):
        mg.capi.roadmap_shortest_paths(3, [[0, 1]], [1.0], [5])


def csr_of(n, edges, weights):
    """Both directions of every edge, neighbours in edge-list order (what the library builds and the _device variant takes)."""
    row = np.zeros(n + 1, np.uint32)
    for a, b in edges.tolist():
        row[a + 1] += 1
        row[b + 1] += 1
    row = np.cumsum(row, dtype=np.uint32)
    cur, col, w = row[:-1].copy(), np.zeros(2 * len(edges), np.uint32), np.zeros(2 * len(edges))
    for (a, b), x in zip(edges.tolist(), weights.tolist()):
        col[cur[a]], w[cur[a]] = b, x
        cur[a] += 1
        col[cur[b]], w[cur[b]] = a, x
        cur[b] += 1
    return row, col, w


def assert_predecessor_tree(src, dist, pred):
    """Every reachable vertex walks to the source in fewer than n hops."""
    n = len(pred)
    for v in range(n):
        if dist[v] == DBL_MAX:
            assert pred[v] == v
            continue
        u, hops = v, 0
        while u != src:
            u, hops = int(pred[u]), hops + 1
            assert hops < n, f"walk from {v} does not reach the source"


def zero_weight_cluster_graph():
    """Source 0 inside a zero-weight triangle (0, 1, 2); a second zero-weight clique (4, 5, 6, 7) one unit away behind vertex 3;
    an edge far too short to register in the sum (8, behind 7); vertex 9 isolated."""
    edges = np.array([[0, 1], [1, 2], [2, 0], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7], [7, 4], [4, 6], [5, 7], [7, 8]], np.uint32)
    w = np.array([0.0, 0.0, 0.0, 1.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1e-30])
    return 10, edges, w


def test_repair_predecessors_breaks_cycles_between_tied_vertices(mg, oracle):
    """The host-side pass of mgodpl_roadmap_shortest_paths (no device involved): predecessors that name each other inside
    clusters joined by vanishing edges — what the per-vertex choice on the device may produce — become a tree; everything
    else is left as it was."""
    n, edges, w = zero_weight_cluster_graph()
    row, col, cw = csr_of(n, edges, w)
    dist, _ = oracle.roadmap_shortest_paths(n, edges, w, [0])
    dist = dist[0]
    assert dist.tolist() == [0.0, 0.0, 0.0, 1.0, 1.5, 1.5, 1.5, 1.5, 1.5, DBL_MAX]
    # worst case the kernel's rule admits: 1 <-> 2 name each other, 4..7 form a ring, 8 hangs off 7
    bad = np.array([0, 2, 1, 2, 5, 6, 7, 4, 7, 9], np.uint32)
    got = mg.capi.roadmap_repair_predecessors(row, col, cw, dist, bad)
    assert got[0] == 0 and got[9] == 9 and got[3] == 2
    assert got[4] == 3                      # the only member with a strictly nearer neighbour on a shortest path
    check_predecessors(n, edges, w, [0], dist[None], got[None])
    assert_predecessor_tree(0, dist, got)
    # an already valid tree is returned unchanged wherever the predecessor is strictly nearer, and stays a tree elsewhere
    good = np.array([0, 0, 0, 2, 3, 4, 4, 4, 7, 9], np.uint32)
    again = mg.capi.roadmap_repair_predecessors(row, col, cw, dist, good)
    assert again[3] == 2 and again[4] == 3
    assert_predecessor_tree(0, dist, again)
    # random graphs with many zero-weight edges, predecessors scrambled inside each tied cluster
    rng = np.random.default_rng(17)
    for _ in range(20):
        n = int(rng.integers(5, 60))
        edges = rng.integers(0, n, (int(rng.integers(n, 4 * n)), 2)).astype(np.uint32)
        w = rng.choice([0.0, 0.0, 1.0, 2.0, 1e-40], len(edges))
        row, col, cw = csr_of(n, edges, w)
        src = int(rng.integers(0, n))
        dist, _ = oracle.roadmap_shortest_paths(n, edges, w, [src])
        dist = dist[0]
        pred = np.arange(n, dtype=np.uint32)
        for v in range(n):  # any neighbour on a shortest path, chosen at random: cycles inside tied clusters are likely
            if v == src or dist[v] == DBL_MAX:
                continue
            cands = [int(col[k]) for k in range(row[v], row[v + 1]) if col[k] != v and dist[col[k]] + cw[k] == dist[v]]
            pred[v] = cands[int(rng.integers(0, len(cands)))]
        got = mg.capi.roadmap_repair_predecessors(row, col, cw, dist, pred)
        check_predecessors(n, edges, w, [src], dist[None], got[None])
        assert_predecessor_tree(src, dist, got)
    with pytest.raises(mg.capi.MgodplError, match="out of range"):
        mg.capi.roadmap_repair_predecessors(row, col, cw, dist, np.full(n, n, np.uint32))


# ---- GPU -------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n,e,g", [(1, 0, 1), (2, 1, 2), (50, 60, 7), (400, 1500, 33), (3000, 12000, 100), (20000, 90000, 64)])
def test_gpu_distances_bit_equal_to_dijkstra(gpu, oracle, n, e, g):
    rng = np.random.default_rng(n + g)
    edges, w = random_graph(rng, n, e)
    src = rng.integers(0, n, g).astype(np.uint32)
    want_d, want_p = oracle.roadmap_shortest_paths(n, edges, w, src, threads=8)
    got = gpu.capi.roadmap_shortest_paths(n, edges, w, src)
    assert np.array_equal(got["dist"], want_d)
    if n <= 3000:
        check_predecessors(n, edges, w, src, got["dist"], got["pred"])
    # continuous random weights: no exact ties, so the predecessor is unique and equals Dijkstra's
    assert np.array_equal(got["pred"], want_p)
    assert np.array_equal(gpu.capi.roadmap_shortest_paths(n, edges, w, src, want_pred=False)["dist"], want_d)


@pytest.mark.gpu
def test_gpu_roadmap_of_a_prm(gpu, oracle):
    """The real thing in small: PRM nodes, exact k-NN, equal_weights_distance weights, goals as sources."""
    orb = oracle.Robot(0.75, (0,))
    nodes = oracle.gen_uniform_states(orb, 4000, seed=3)
    idx, dist = gpu.capi.knn(nodes, 6, causal=True)
    keep = idx >= 0
    edges = np.stack([idx[keep].astype(np.uint32), np.repeat(np.arange(len(nodes), dtype=np.uint32), 6)[keep.reshape(-1)]], axis=1)
    w = dist[keep]
    src = np.arange(0, 4000, 31, dtype=np.uint32)
    want_d, want_p = oracle.roadmap_shortest_paths(len(nodes), edges, w, src, threads=8)
    got = gpu.capi.roadmap_shortest_paths(len(nodes), edges, w, src)
    assert np.array_equal(got["dist"], want_d) and np.array_equal(got["pred"], want_p)
    assert (want_d < DBL_MAX).all()  # causal k-NN graphs are connected
    # symmetric: d(a -> b) along the reversed path sums the same weights in the other order; equal to ~1 ulp per hop only
    sub = got["dist"][:, src]
    assert np.allclose(sub, sub.T, rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_gpu_ties_zero_weights_and_long_chains(gpu, oracle):
    # a 3000-vertex chain: 2999 hops from one end, far more sweeps than a roadmap ever needs
    n = 3000
    edges = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1).astype(np.uint32)
    w = np.full(n - 1, 0.1)
    got = gpu.capi.roadmap_shortest_paths(n, edges, w, [0, n - 1, n // 2])
    want_d, want_p = oracle.roadmap_shortest_paths(n, edges, w, [0, n - 1, n // 2])
    assert np.array_equal(got["dist"], want_d) and np.array_equal(got["pred"], want_p)
    assert got["sweeps"] >= 8
    # integer weights on a grid: many exact ties; distances still equal, predecessors valid (tie rule documented in the header)
    k = 24
    idx = np.arange(k * k).reshape(k, k)
    edges = np.concatenate([np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1), np.stack([idx[:-1].ravel(), idx[1:].ravel()], 1)]).astype(np.uint32)
    rng = np.random.default_rng(5)
    w = rng.integers(1, 4, len(edges)).astype(np.float64)
    src = rng.integers(0, k * k, 40).astype(np.uint32)
    got = gpu.capi.roadmap_shortest_paths(k * k, edges, w, src)
    want_d, _ = oracle.roadmap_shortest_paths(k * k, edges, w, src)
    assert np.array_equal(got["dist"], want_d)
    check_predecessors(k * k, edges, w, src, got["dist"], got["pred"])
    # self-loops, parallel edges, a zero-weight edge, duplicate sources, an isolated vertex
    edges = np.array([[0, 0], [0, 1], [0, 1], [1, 2], [2, 3], [3, 3]], np.uint32)
    w = np.array([5.0, 2.0, 1.5, 0.0, 1.0, 0.5])
    got = gpu.capi.roadmap_shortest_paths(5, edges, w, [0, 0, 3])
    want_d, _ = oracle.roadmap_shortest_paths(5, edges, w, [0, 0, 3])
    assert np.array_equal(got["dist"], want_d) and got["dist"][0].tolist() == [0.0, 1.5, 1.5, 2.5, DBL_MAX]
    assert got["pred"][0].tolist()[:2] == [0, 0] and got["pred"][0][4] == 4 and got["pred"][0][3] == 2
    # zero-weight clusters, one of them around the source, and an edge too short to register: distances exact, predecessors a tree
    cn, cedges, cw = zero_weight_cluster_graph()
    got = gpu.capi.roadmap_shortest_paths(cn, cedges, cw, [0, 5, 8])
    want_d, _ = oracle.roadmap_shortest_paths(cn, cedges, cw, [0, 5, 8])
    assert np.array_equal(got["dist"], want_d)
    check_predecessors(cn, cedges, cw, [0, 5, 8], got["dist"], got["pred"])
    for i, src in enumerate([0, 5, 8]):
        assert_predecessor_tree(src, got["dist"][i], got["pred"][i])
    # no sources / no vertices
    assert gpu.capi.roadmap_shortest_paths(5, edges, w, [])["dist"].shape == (0, 5)


@pytest.mark.gpu
def test_gpu_roadmap_full_size_properties(gpu):
    """BASELINE-sized roadmap (100 k nodes, ~1 M edges): checked through the fixed-point property, which is size independent:
    every edge satisfies d[v] <= d[u] + w, and every reached vertex has a neighbour attaining equality (its predecessor)."""
    rng = np.random.default_rng(11)
    n, k = 100_000, 10
    pts = rng.uniform(-5, 5, (n, 3))
    # cheap spatial neighbours: sort by a coarse cell key and connect to the next k in that order, plus random long links
    order = np.lexsort((pts[:, 2] // 1.0, pts[:, 1] // 1.0, pts[:, 0] // 1.0))
    a = np.repeat(order[:-k], k)
    b = np.concatenate([order[i + 1:i + 1 + k] for i in range(0, n - k)]) if n <= 2000 else np.stack([order[j:n - k + j] for j in range(1, k + 1)], 1).ravel()
    edges = np.stack([a, b], 1).astype(np.uint32)
    w = np.linalg.norm(pts[edges[:, 0]] - pts[edges[:, 1]], axis=1) + 1e-3
    src = rng.integers(0, n, 96).astype(np.uint32)
    got = gpu.capi.roadmap_shortest_paths(n, edges, w, src)
    d, p = got["dist"], got["pred"]
    assert (d < DBL_MAX).all() and (d[np.arange(len(src)), src] == 0).all()
    du, dv = d[:, edges[:, 0]], d[:, edges[:, 1]]
    assert (dv <= du + w).all() and (du <= dv + w).all()
    # predecessor attains equality through some edge between the two vertices
    key = {}
    for (x, y), ww in zip(edges.tolist(), w.tolist()):
        key.setdefault((x, y), []).append(ww)
        key.setdefault((y, x), []).append(ww)
    for s in (0, 47, 95):
        vs = rng.integers(0, n, 2000)
        for v in vs.tolist():
            if v == src[s]:
                continue
            u = int(p[s, v])
            assert any(d[s, u] + ww == d[s, v] for ww in key[(u, v)])
