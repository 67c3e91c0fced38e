"""Tree-model ingest on the host: mgodpl::from_dae (host/mgodpl/tree_meshes.hpp) against an independent
assimp-style reading of the same COLLADA documents (tests/dae_reader.py; mesh_from_dae.cpp:17-85).  assimp itself is absent,
so what is pinned here is assimp's documented behaviour — aiMesh order = first instantiation in the node hierarchy, one
aiMesh per primitive block, identical positions joined in first-seen order, float32 positions — on hand-made documents and
on the one real model the reference ships (test_robots/meshes/drone.dae, read in place when /root/reference is present)."""
import json
import os
import subprocess

import numpy as np
import pytest

import dae_reader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRONE = "/root/reference/test_robots/meshes/drone.dae"


@pytest.fixture(scope="module")
def dae_summary(mg):
    mg.capi.lib()  # the tool links the in-tree library (host code only; no device needed)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    exe = os.path.join(ROOT, "tests", "cpp", "dae_summary")

    def run(path):
        p = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=60)
        return p.returncode, json.loads(p.stdout)
    return run


def same(got, want):
    assert got["n_vertices"] == want["n_vertices"] and got["n_triangles"] == want["n_triangles"]
    assert got["triangle_checksum"] == want["triangle_checksum"]
    assert got["vertex_checksum"] == want["vertex_checksum"]      # bit-equal: same float32 -> fp64 conversion, same order
    assert got["lo"] == want["lo"] and got["hi"] == want["hi"]


def geometry(gid, positions, prim):
    flat = " ".join(repr(float(x)) for p in positions for x in p)
    return (f'<geometry id="{gid}"><mesh><source id="{gid}-pos"><float_array id="{gid}-arr" count="{3 * len(positions)}">{flat}</float_array>'
            f'<technique_common><accessor source="#{gid}-arr" count="{len(positions)}" stride="3"><param name="X" type="float"/>'
            f'<param name="Y" type="float"/><param name="Z" type="float"/></accessor></technique_common></source>'
            f'<source id="{gid}-nrm"><float_array id="{gid}-narr" count="3">0 0 1</float_array><technique_common>'
            f'<accessor source="#{gid}-narr" count="1" stride="3"/></technique_common></source>'
            f'<vertices id="{gid}-vtx"><input semantic="POSITION" source="#{gid}-pos"/></vertices>{prim}</mesh></geometry>')


def document(geometries, scene):
    return ('<?xml version="1.0" encoding="utf-8"?>\n<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">'
            '<asset><unit name="inch" meter="0.0254"/><up_axis>Y_UP</up_axis></asset><!-- a comment -->'
            f'<library_geometries>{"".join(geometries)}</library_geometries>{scene}</COLLADA>\n')


QUAD = [(0, 0, 0), (10, 0, 0), (10, 10, 0), (0, 10, 0), (10, 0, 0)]       # position 4 repeats position 1: joined
GEOMS = [
    # a: <triangles>, VERTEX at offset 0 and a NORMAL input at offset 1; corner order makes position 2 the first one seen
    geometry("a", QUAD, '<triangles count="2"><input semantic="VERTEX" source="#a-vtx" offset="0"/><input semantic="NORMAL" source="#a-nrm" offset="1"/>'
             '<p>2 0 1 0 0 0  0 0 4 0 2 0</p></triangles>'),
    # b: <polylist> with a quad and a pentagon (fan-triangulated), VERTEX at offset 1
    geometry("b", [(1, 2, 3), (4, 2, 3), (4, 6, 3), (1, 6, 3), (0, 4, 3)],
             '<polylist count="2"><input semantic="NORMAL" source="#b-nrm" offset="0"/><input semantic="VERTEX" source="#b-vtx" offset="1"/>'
             '<vcount>4 5</vcount><p>0 0 0 1 0 2 0 3   0 0 0 1 0 2 0 3 0 4</p></polylist>'),
    # c: two primitive blocks over the same positions (two aiMeshes: joined separately), plus lines that are dropped
    geometry("c", [(-1, -1, -1), (1, -1, -1), (0, 1, -1), (0, 0, 5)],
             '<lines count="1"><input semantic="VERTEX" source="#c-vtx" offset="0"/><p>0 1</p></lines>'
             '<triangles count="1"><input semantic="VERTEX" source="#c-vtx" offset="0"/><p>0 1 2</p></triangles>'
             '<polygons count="1"><input semantic="VERTEX" source="#c-vtx" offset="0"/><p>0 1 3</p></polygons>'),
    # d: never instantiated: assimp does not load it
    geometry("d", [(7, 7, 7), (8, 7, 7), (7, 8, 7)], '<triangles count="1"><input semantic="VERTEX" source="#d-vtx" offset="0"/><p>0 1 2</p></triangles>'),
]
# node order: c first (own instance before the children), then b (child), a through an <instance_node>, c again (no second copy)
SCENE = ('<library_nodes><node id="lib"><instance_geometry url="#a"/></node></library_nodes>'
         '<library_visual_scenes><visual_scene id="other"><node id="x"><instance_geometry url="#d"/></node></visual_scene>'
         '<visual_scene id="main"><node id="n1"><matrix>2 0 0 5 0 2 0 5 0 0 2 5 0 0 0 1</matrix><node id="n2"><instance_geometry url="#b"/></node>'
         '<instance_geometry url="#c"/></node><node id="n3"><instance_node url="#lib"/><instance_geometry url="#c"/></node></visual_scene>'
         '</library_visual_scenes><scene><instance_visual_scene url="#main"/></scene>')


def test_from_dae_follows_assimps_mesh_order_and_vertex_joining(dae_summary, tmp_path):
    f = tmp_path / "model.dae"       # not a tree file name: coordinates / 10
    f.write_text(document(GEOMS, SCENE))
    v, t = dae_reader.read(f)
    # the expectation, spelled out: c's two blocks (3 + 3 vertices), b (5), a (3 after joining) — d absent
    assert len(v) == 3 + 3 + 5 + 3 and len(t) == 1 + 1 + (2 + 3) + 2
    assert v[0].tolist() == [-0.1, -0.1, -0.1] and v[5].tolist() == [0.0, 0.0, 0.5]
    assert t[:2].tolist() == [[0, 1, 2], [3, 4, 5]]
    assert t[2:7].tolist() == [[6, 7, 8], [6, 8, 9], [6, 7, 8], [6, 8, 9], [6, 9, 10]]
    assert v[11].tolist() == [1.0, 1.0, 0.0]                      # a: position 2 was seen first
    assert t[7:].tolist() == [[11, 12, 13], [13, 12, 11]]         # ... and position 4 is position 1 again (vertex 12)
    rc, got = dae_summary(f)
    assert rc == 0
    same(got, dae_reader.summary(v, t))


def test_from_dae_tree_file_names_convert_inches_and_swap_axes(dae_summary, tmp_path):
    f = tmp_path / "appletree_trunk.dae"
    f.write_text(document(GEOMS[1:2], '<library_visual_scenes><visual_scene id="s"><node id="n"><instance_geometry url="#b"/></node></visual_scene>'
                                      '</library_visual_scenes><scene><instance_visual_scene url="#s"/></scene>'))
    v, t = dae_reader.read(f)
    s = 100.0 / 2.54
    assert v[0].tolist() == [1.0 / s, 3.0 / s, 2.0 / s]           # mesh_from_dae.cpp:55-63: y and z swapped
    rc, got = dae_summary(f)
    assert rc == 0
    same(got, dae_reader.summary(v, t))


def test_from_dae_without_a_visual_scene_reads_in_library_order(dae_summary, tmp_path):
    f = tmp_path / "loose.dae"
    f.write_text(document(GEOMS, ""))
    v, t = dae_reader.read(f)
    assert len(t) == 2 + 5 + 2 + 1                                # a, b, c, d
    rc, got = dae_summary(f)
    assert rc == 0
    same(got, dae_reader.summary(v, t))


def test_from_dae_errors(dae_summary, tmp_path):
    rc, got = dae_summary(tmp_path / "missing.dae")
    assert rc == 1 and "Failed to load mesh from" in got["error"] and "Are you in the right working directory" in got["error"]
    f = tmp_path / "broken.dae"
    f.write_text("<COLLADA><library_geometries>")
    rc, got = dae_summary(f)
    assert rc == 1 and "malformed" in got["error"]
    f.write_text(document([geometry("a", QUAD, '<triangles count="1"><input semantic="VERTEX" source="#a-vtx" offset="0"/><p>0 1 9</p></triangles>')],
                          '<library_visual_scenes><visual_scene id="s"><node id="n"><instance_geometry url="#a"/></node></visual_scene></library_visual_scenes>'))
    rc, got = dae_summary(f)
    assert rc == 1 and "index out of range" in got["error"]


def test_golden_summary_of_the_reference_drone_model_is_consistent():
    """The committed summary (tests/golden/make_golden_dae.py): 4 rotor discs of 16 triangles, 352 + 1728 in the two bodies."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "drone_dae_summary.json")))
    assert g["n_triangles"] == 4 * 16 + 352 + 1728 and g["n_vertices"] == 1114
    assert np.allclose(g["lo"][:2], [-0.50676, -0.50676], atol=1e-5) and np.allclose(g["hi"][:2], [0.50676, 0.50676], atol=1e-5)


@pytest.mark.skipif(not os.path.exists(DRONE), reason="the reference checkout (and its drone.dae) is only present in the build container")
def test_from_dae_reads_the_reference_drone_model(dae_summary):
    """A real exporter's document (Blender 2.93: six geometries, VERTEX + NORMAL inputs, node matrices that the reference ignores)."""
    rc, got = dae_summary(DRONE)
    assert rc == 0
    same(got, dae_reader.summary(*dae_reader.read(DRONE)))
    same(got, json.load(open(os.path.join(ROOT, "tests", "golden", "drone_dae_summary.json"))))


def test_from_dae_survives_damaged_documents(dae_summary, tmp_path):
    """Truncated, corrupted and spliced documents end in a mesh or in an exception — never in a crash (exit codes 0 / 1 only)."""
    import random
    doc = document(GEOMS, SCENE).encode()
    rnd = random.Random(1)
    splices = [b"<p>1 2 99999999999999999999</p>", b"<node>", b"</mesh>", b"-5 ", b"<![CDATA[", b"<!--", b'<instance_node url="#n1"/>', b"1e400 ", b"nan "]
    f = tmp_path / "damaged.dae"
    for i in range(120):
        b = bytearray(doc)
        if i % 4 == 0:
            b = b[:rnd.randrange(len(b))]
        elif i % 4 == 1:
            for _ in range(rnd.randrange(1, 6)):
                b[rnd.randrange(len(b))] = rnd.choice(b'<>/"= 0-9ae\n')
        elif i % 4 == 2:
            a = rnd.randrange(len(b))
            del b[a:rnd.randrange(a, min(len(b), a + 200))]
        else:
            a = rnd.randrange(len(b))
            b[a:a] = rnd.choice(splices)
        f.write_bytes(bytes(b))
        rc, out = dae_summary(f)
        assert rc in (0, 1) and (("error" in out) == (rc == 1)), (i, rc, out)
