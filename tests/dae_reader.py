"""An independent reading of a COLLADA file, written from assimp's documented behaviour rather than from the product's
reader: test infrastructure for tests/test_ingest.py and tests/golden/make_golden_dae.py.

What mesh_from_dae.cpp:17-85 gets from assimp (aiProcess_Triangulate | JoinIdenticalVertices | SortByPType |
RemoveComponent with everything but positions removed) and then does with it:
  * one aiMesh per primitive block of a geometry, created the first time a node of the instantiated visual scene refers to
    that geometry (depth first, a node's own instances before its children); geometries nobody instantiates are absent;
  * its vertices are the corner positions in first-seen order with identical positions joined (float32, aiVector3D);
  * the reference concatenates the aiMeshes, ignores every node transform, and converts units by file name:
    tree meshes (x, z, y) * 2.54 / 100, everything else / 10.
"""
import xml.etree.ElementTree as ET

import numpy as np

NS = "{http://www.collada.org/2005/11/COLLADASchema}"


def _strip(ref):
    return ref[1:] if ref.startswith("#") else ref


def read(path):
    root = ET.parse(path).getroot()
    geometries = {g.get("id"): g for g in root.iter(NS + "geometry")}
    nodes = {n.get("id"): n for n in root.iter(NS + "node")}
    scenes = {s.get("id"): s for s in root.iter(NS + "visual_scene")}
    start = None
    inst = root.find(f"{NS}scene/{NS}instance_visual_scene")
    if inst is not None:
        start = scenes.get(_strip(inst.get("url", "")))
    if start is None and scenes:
        start = next(iter(scenes.values()))
    order = []

    def walk(node, active):
        if node in active:
            return
        for ig in node.findall(NS + "instance_geometry"):
            gid = _strip(ig.get("url", ""))
            if gid in geometries and gid not in order:
                order.append(gid)
        for child in node:
            if child.tag == NS + "node":
                walk(child, active + [node])
            elif child.tag == NS + "instance_node" and _strip(child.get("url", "")) in nodes:
                walk(nodes[_strip(child.get("url", ""))], active + [node])

    if start is not None:
        for n in start.findall(NS + "node"):
            walk(n, [])
    else:
        order = list(geometries)
    name = str(path)
    tree = any(w in name for w in ("apples", "fruit", "leaves", "trunk"))
    vertices, triangles = [], []
    for gid in order:
        mesh = geometries[gid].find(NS + "mesh")
        if mesh is None:
            continue
        sources = {}
        for src in mesh.findall(NS + "source"):
            arr = src.find(NS + "float_array")
            if arr is None:
                continue
            acc = src.find(f"{NS}technique_common/{NS}accessor")
            stride = int(acc.get("stride", "3")) if acc is not None else 3
            vals = np.array((arr.text or "").split(), dtype=np.float64).astype(np.float32)
            if stride >= 3:
                sources[src.get("id")] = vals[: len(vals) // stride * stride].reshape(-1, stride)[:, :3]
        position_of = {}
        for v in mesh.findall(NS + "vertices"):
            for inp in v.findall(NS + "input"):
                if inp.get("semantic") == "POSITION":
                    position_of[v.get("id")] = _strip(inp.get("source"))
        for prim in mesh:
            kind = prim.tag[len(NS):]
            if kind not in ("triangles", "polylist", "polygons"):
                continue
            inputs = prim.findall(NS + "input")
            if not inputs:
                continue
            n_inputs = max(int(i.get("offset", "0")) for i in inputs) + 1
            vin = [i for i in inputs if i.get("semantic") == "VERTEX"]
            if not vin or position_of.get(_strip(vin[0].get("source"))) not in sources:
                continue
            pos = sources[position_of[_strip(vin[0].get("source"))]]
            voff = int(vin[0].get("offset", "0"))
            polys = []
            if kind == "polygons":
                for p in prim.findall(NS + "p"):
                    idx = np.array(p.text.split(), dtype=np.int64).reshape(-1, n_inputs)[:, voff]
                    polys.append(idx)
            else:
                idx = np.array((prim.find(NS + "p").text or "").split(), dtype=np.int64)
                idx = idx[: len(idx) // n_inputs * n_inputs].reshape(-1, n_inputs)[:, voff]
                if kind == "triangles":
                    polys = [idx[k:k + 3] for k in range(0, len(idx) - 2, 3)]
                else:
                    k = 0
                    for c in (int(x) for x in prim.find(NS + "vcount").text.split()):
                        polys.append(idx[k:k + c])
                        k += c
            seen = {}

            def corner(pi):
                key = pos[pi].tobytes()
                if key not in seen:
                    seen[key] = len(vertices)
                    x, y, z = (float(c) for c in pos[pi])
                    if tree:
                        s = 100.0 / 2.54
                        vertices.append((x / s, z / s, y / s))
                    else:
                        vertices.append((x / 10.0, y / 10.0, z / 10.0))
                return seen[key]

            for poly in polys:
                for k in range(1, len(poly) - 1):
                    # evaluation order a, b, c matters: it decides which corner is "first seen"
                    a = corner(poly[0])
                    b = corner(poly[k])
                    c = corner(poly[k + 1])
                    triangles.append((a, b, c))
    return np.array(vertices, dtype=np.float64).reshape(-1, 3), np.array(triangles, dtype=np.int64).reshape(-1, 3)


def summary(vertices, triangles):
    """The same figures tests/cpp/dae_summary.cpp prints (sequential fp64 sums, so the order of operations is the same)."""
    vsum = 0.0
    for i, (x, y, z) in enumerate(vertices.tolist()):
        vsum += float(i % 97 + 1) * (x + 2.0 * y + 3.0 * z)
    tsum = 0
    for t, (a, b, c) in enumerate(triangles.tolist()):
        tsum += (t % 89 + 1) * (a + 3 * b + 7 * c)
    lo = vertices.min(axis=0).tolist() if len(vertices) else [float("inf")] * 3
    hi = vertices.max(axis=0).tolist() if len(vertices) else [float("-inf")] * 3
    return dict(n_vertices=len(vertices), n_triangles=len(triangles), lo=lo, hi=hi, vertex_checksum=vsum,
                triangle_checksum=tsum % (1 << 64))
