"""Dev tool: BVH quality (SAH cost, node visits) and bench-workload time with / without the rotation refinement."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE); mesh = tree.collision_mesh(True)
model, a, b = bench.make_inputs(mg, 0); robot = model.to_device()
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
m = len(a)
d_hit = torch.zeros((m + 31)//32, dtype=torch.int32, device=dev); d_toi = torch.zeros(m, dtype=torch.float64, device=dev)
d_steps = torch.zeros(m, dtype=torch.int32, device=dev)
ref = None
for refine in (False, True):
    for leaf in (1, 2, 4):
        scene = mg.Scene.from_mesh(mesh, counters=True, max_leaf_size=leaf, refine=refine)
        info = scene.info()
        for _ in range(3):
            mg.capi.check_motions_device(scene, robot, d_a, d_b, m, 9, 2, d_hit, 0.1, d_toi, d_steps, stream=stream)
        torch.cuda.synchronize(); scene.counters(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            mg.capi.check_motions_device(scene, robot, d_a, d_b, m, 9, 2, d_hit, 0.1, d_toi, d_steps, stream=stream)
        e1.record(); torch.cuda.synchronize()
        c = scene.counters()
        h = d_hit.cpu().numpy().copy()
        if ref is None: ref = h
        print(f"refine={refine} leaf={leaf}: build {info.build_ms:.2f} ms nodes={info.n_nodes} depth={info.max_depth} sah={info.sah_cost:.1f} "
              f"nodes/box={c['nodes_visited']/max(c['boxes_traversed'],1):.1f} leaf/box={c['leaf_tests']/max(c['boxes_traversed'],1):.2f} "
              f"ms={e0.elapsed_time(e1)/10:.3f} same_result={np.array_equal(h, ref)}")
