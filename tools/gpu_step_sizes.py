"""Dev tool: the motion step at several batch sizes (latency floor of the pipeline), headline scene (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE)
scene = mg.Scene.from_mesh(tree.collision_mesh(True))
model, a, b = bench.make_inputs(mg, 0)
robot = model.to_device()
dev = torch.device("cuda")
d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
M = len(a)
d_hit = torch.zeros((M + 31) // 32, dtype=torch.int32, device=dev)
d_toi = torch.zeros(M, dtype=torch.float64, device=dev)
d_steps = torch.zeros(M, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
sizes = [int(x) for x in sys.argv[1:]] or [1000, 5000, 25000, 50000, 100000]
for m in sizes:
    ts = []
    for k in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mg.capi.check_motions_device(scene, robot, d_a, d_b, m, 9, 2, d_hit, 0.1, d_toi, d_steps, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"m={m}: step ms (warm L2) min {min(ts[1:]):.4f} median {sorted(ts[1:])[2]:.4f}", flush=True)
