"""Procedural fruit trees and orchard rows (synthetic stand-ins for the reference's 3d-models, which are an
un-fetched git-LFS blob — SURVEY.md "facts", Appendix C).

Output conventions follow what the reference's loader produces (src/experiment_utils/mesh_from_dae.cpp:55-62,
TreeMeshes.h:19-24): metres, z-up, trunk rooted at the world origin, a *trunk* mesh (trunk + branches, the set the
reference actually collides with — fcl_utils.cpp:58-60), a separate *leaves* mesh, and fruit centres
(TreeMeshes.cpp:208-214).  Orchard rows use makeSingleRowOrchard's layout (TreeMeshes.cpp:197-206): n trees, 2 m
apart along x, first at x = -n.

Everything is deterministic in ``seed`` (numpy PCG64).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Mesh:
    """Indexed triangle soup, as src/planning/Mesh.h:19-22 (vertices fp64, triangles uint32 here)."""
    vertices: np.ndarray  # (nv, 3) float64
    triangles: np.ndarray  # (nt, 3) uint32

    @property
    def n_triangles(self) -> int:
        return len(self.triangles)

    def aabb(self):
        return self.vertices.min(axis=0), self.vertices.max(axis=0)

    def translated(self, offset) -> "Mesh":
        return Mesh(self.vertices + np.asarray(offset, dtype=np.float64), self.triangles.copy())


def merge(meshes) -> Mesh:
    verts, tris, base = [], [], 0
    for m in meshes:
        verts.append(m.vertices)
        tris.append(m.triangles.astype(np.int64) + base)
        base += len(m.vertices)
    if not verts:
        return Mesh(np.zeros((0, 3)), np.zeros((0, 3), np.uint32))
    return Mesh(np.concatenate(verts), np.concatenate(tris).astype(np.uint32))


@dataclass
class TreeMeshes:
    """Mirror of mgodpl::tree_meshes::TreeMeshes (src/experiment_utils/TreeMeshes.h:19-24) with fruit reduced to
    their centres (computeFruitPositions, TreeMeshes.cpp:208-214)."""
    tree_name: str
    trunk_mesh: Mesh
    leaves_mesh: Mesh
    fruit_centers: np.ndarray  # (F, 3)

    def collision_mesh(self, with_leaves: bool) -> Mesh:
        """'T' scene = trunk only (the reference's set); 'T+L' = trunk + leaves (north_star extension)."""
        return merge([self.trunk_mesh, self.leaves_mesh]) if with_leaves else self.trunk_mesh


def _frame(d):
    """Two unit vectors orthogonal to unit direction d."""
    a = np.array([1.0, 0, 0]) if abs(d[0]) < 0.9 else np.array([0, 1.0, 0])
    u = np.cross(d, a)
    u /= np.linalg.norm(u)
    return u, np.cross(d, u)


def _tube(points, radii, nside):
    """Triangulated tapered tube along a polyline, closed with an end cap."""
    n = len(points)
    d = np.gradient(points, axis=0)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    u, v = _frame(d[0])
    ang = np.arange(nside) * (2 * np.pi / nside)
    rings = []
    for i in range(n):
        # transport the frame along the polyline
        u = u - d[i] * np.dot(u, d[i])
        u /= np.linalg.norm(u)
        v = np.cross(d[i], u)
        rings.append(points[i] + radii[i] * (np.cos(ang)[:, None] * u + np.sin(ang)[:, None] * v))
    verts = np.concatenate(rings + [points[-1:]])
    i = np.arange(n - 1)[:, None] * nside
    j = np.arange(nside)[None, :]
    jn = (j + 1) % nside
    a, b, c, e = i + j, i + jn, i + nside + j, i + nside + jn
    tris = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([b, e, c], -1).reshape(-1, 3)])
    tip = n * nside
    cap = np.stack([(n - 1) * nside + j[0], (n - 1) * nside + jn[0], np.full(nside, tip)], -1)
    return verts, np.concatenate([tris, cap])


def make_tree(seed: int = 42, trunk_triangles: int = 50_000, leaf_triangles: int = 150_000, n_fruit: int = 500,
              name: str | None = None) -> TreeMeshes:
    """One synthetic fruit tree: ~3 m tall, canopy radius ~1.5 m, canopy centre near (0, 0, 2.1)."""
    rng = np.random.default_rng(seed)
    # ---- skeleton: (start, direction, length, r0, r1, level) ------------------------------------------------
    fan = [6, 4, 4, 3]
    branches = [(np.zeros(3), np.array([0.0, 0.0, 1.0]), 1.35, 0.13, 0.085, 0)]
    frontier = [branches[0]]
    for level, nchild in enumerate(fan, start=1):
        nxt = []
        for (p, d, L, r0, r1, _) in frontier:
            u, v = _frame(d)
            for k in range(nchild):
                s = rng.uniform(0.55, 1.0) if level > 1 else rng.uniform(0.75, 1.0)
                start = p + d * (L * s)
                az = 2 * np.pi * (k + rng.uniform(-0.3, 0.3)) / nchild
                spread = rng.uniform(0.55, 1.05)
                nd = d * np.cos(spread) + (u * np.cos(az) + v * np.sin(az)) * np.sin(spread)
                nd[2] += 0.25 * (1.0 - level / 5)  # gentle upward bias
                nd /= np.linalg.norm(nd)
                nl = L * rng.uniform(0.55, 0.8)
                nr0 = r1 * (0.75 if level == 1 else 0.6)
                b = (start, nd, nl, nr0, nr0 * 0.55, level)
                branches.append(b)
                nxt.append(b)
        frontier = nxt
    # ---- trunk/branch tubes ----------------------------------------------------------------------------------
    nb = len(branches)
    per_branch = max(trunk_triangles / nb, 16.0)
    verts, tris, base = [], [], 0
    twig_pts, twig_dirs = [], []
    for (p, d, L, r0, r1, level) in branches:
        nside = 8 if level == 0 else (6 if level <= 2 else 5)
        w = 4.0 if level == 0 else (1.6 if level == 1 else 1.0)
        nseg = max(2, int(round((per_branch * w - nside) / (2 * nside))))
        t = np.linspace(0, 1, nseg + 1)
        u, v = _frame(d)
        bend = 0.08 * L * np.sin(np.pi * t) * rng.uniform(-1, 1)
        pts = p + np.outer(t * L, d) + np.outer(bend, u) + np.outer(0.5 * bend, v)
        vv, tt = _tube(pts, r0 + (r1 - r0) * t, nside)
        verts.append(vv)
        tris.append(tt + base)
        base += len(vv)
        if level >= 3:
            twig_pts.append(pts)
            twig_dirs.append(np.repeat(d[None], len(pts), 0))
    trunk = Mesh(np.concatenate(verts), np.concatenate(tris).astype(np.uint32))
    twig_pts = np.concatenate(twig_pts)
    twig_dirs = np.concatenate(twig_dirs)
    # ---- leaves: small quads around twigs ---------------------------------------------------------------------
    n_leaves = leaf_triangles // 2
    if n_leaves:
        at = rng.integers(0, len(twig_pts), n_leaves)
        c = twig_pts[at] + rng.normal(0, 0.07, (n_leaves, 3))
        a = rng.normal(size=(n_leaves, 3))
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        b = np.cross(a, rng.normal(size=(n_leaves, 3)))
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        ha = a * (rng.uniform(0.025, 0.045, n_leaves))[:, None]
        hb = b * (rng.uniform(0.012, 0.022, n_leaves))[:, None]
        lv = np.stack([c - ha - hb, c + ha - hb, c + ha + hb, c - ha + hb], 1).reshape(-1, 3)
        q = np.arange(n_leaves, dtype=np.int64)[:, None] * 4
        lt = np.concatenate([q + np.array([[0, 1, 2]]), q + np.array([[0, 2, 3]])], 1).reshape(-1, 3)
        leaves = Mesh(lv, lt.astype(np.uint32))
    else:
        leaves = Mesh(np.zeros((0, 3)), np.zeros((0, 3), np.uint32))
    # ---- fruit: hang just below random twig points -------------------------------------------------------------
    at = rng.integers(0, len(twig_pts), n_fruit)
    fruit = twig_pts[at] + rng.normal(0, 0.03, (n_fruit, 3)) - np.array([0, 0, 0.06])
    return TreeMeshes(name or f"synthetic_tree_{seed}", trunk, leaves, fruit)


def make_single_row_orchard(trees) -> TreeMeshes:
    """makeSingleRowOrchard (TreeMeshes.cpp:197-206) flattened into one TreeMeshes: tree i at x = -n + 2 i."""
    n = len(trees)
    x0 = n * 2.0 * -0.5
    off = [np.array([x0 + 2.0 * i, 0.0, 0.0]) for i in range(n)]
    return TreeMeshes(
        f"orchard_row_{n}",
        merge([t.trunk_mesh.translated(o) for t, o in zip(trees, off)]),
        merge([t.leaves_mesh.translated(o) for t, o in zip(trees, off)]),
        np.concatenate([t.fruit_centers + o for t, o in zip(trees, off)]),
    )


def make_orchard(n_trees: int = 8, seed: int = 42, trunk_triangles: int = 62_500, leaf_triangles: int = 187_500,
                 n_fruit: int = 500) -> TreeMeshes:
    """BASELINE config 5: 8 distinct synthetic trees, ~2 M triangles in total by default."""
    return make_single_row_orchard(
        [make_tree(seed + i, trunk_triangles, leaf_triangles, n_fruit) for i in range(n_trees)])


def single_triangle(p0, p1, p2) -> Mesh:
    return Mesh(np.array([p0, p1, p2], dtype=np.float64), np.array([[0, 1, 2]], np.uint32))
