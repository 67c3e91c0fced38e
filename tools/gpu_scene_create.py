"""Dev tool: wall time of scene creation (upload + device build + host-side 4-wide fold) for the tree and the orchard."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mgodpl_b200 as mg
import bench
for name, mesh in (("tree 200k", mg.scenes.make_tree(**bench.TREE).collision_mesh(True)), ("orchard 2M", mg.scenes.make_orchard().collision_mesh(True))):
    for rep in range(3):
        t0 = time.perf_counter(); scene = mg.Scene.from_mesh(mesh); dt = (time.perf_counter() - t0) * 1e3
        info = scene.info()
    print("%s: create %.1f ms wall (device build %.1f ms), %d binary + %d wide records (MGODPL_WIDE_BVH=%s)" % (name, dt, info.build_ms, info.n_nodes, info.n_wide_nodes, os.environ.get("MGODPL_WIDE_BVH", "1")))
