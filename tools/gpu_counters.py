"""Dev tool: device-side work counters of the bench workload for several BVH leaf sizes (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mgodpl_b200 as mg
import bench

tree = mg.scenes.make_tree(**bench.TREE)
mesh = tree.collision_mesh(True)
model, a, b = bench.make_inputs(mg, 0)
robot = model.to_device()
dev = torch.device("cuda")
d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
m = len(a)
d_hit = torch.zeros((m + 31) // 32, dtype=torch.int32, device=dev)
d_toi = torch.zeros(m, dtype=torch.float64, device=dev)
d_steps = torch.zeros(m, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for leaf in [int(x) for x in (sys.argv[1:] or [1, 2, 4, 8])]:
    for counters in (True, False):
        scene = mg.Scene.from_mesh(mesh, counters=counters, max_leaf_size=leaf)
        info = scene.info()
        for _ in range(3):
            mg.capi.check_motions_device(scene, robot, d_a, d_b, m, 9, 2, d_hit, 0.1, d_toi, d_steps, stream=stream)
        torch.cuda.synchronize()
        if counters:
            scene.counters(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            mg.capi.check_motions_device(scene, robot, d_a, d_b, m, 9, 2, d_hit, 0.1, d_toi, d_steps, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if counters:
            c = scene.counters()
            k = {x: v / 5 for x, v in c.items()}
            print(f"leaf={leaf} nodes={info.n_nodes} depth={info.max_depth} sah={info.sah_cost:.1f} | states={k['states_checked']:.3g} "
                  f"boxes={k['boxes_traversed']:.3g} nodes/box={k['nodes_visited']/max(k['boxes_traversed'],1):.1f} "
                  f"leaftests/box={k['leaf_tests']/max(k['boxes_traversed'],1):.2f} max_records/box={c['max_nodes_per_box']} ms(with counters)={ms:.3f}")
        else:
            print(f"leaf={leaf} ms={ms:.3f}")
