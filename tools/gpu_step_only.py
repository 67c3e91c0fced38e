"""Dev tool: just the headline step (100k motions, 200k-triangle tree), a few times — for ncu captures and launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
tree = mg.scenes.make_tree(**bench.TREE)
scene = mg.Scene.from_mesh(tree.collision_mesh(True))
model, a, b = bench.make_inputs(mg, 0)
robot = model.to_device()
dev = torch.device("cuda")
d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
m = len(a)
d_hit = torch.zeros((m + 31) // 32, dtype=torch.int32, device=dev)
d_toi = torch.zeros(m, dtype=torch.float64, device=dev)
d_steps = torch.zeros(m, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
ts = []
for k in range(n_steps):
    flush.fill_(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mg.capi.check_motions_device(scene, robot, d_a, d_b, m, 9, 2, d_hit, 0.1, d_toi, d_steps, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("step ms:", " ".join(f"{t:.4f}" for t in ts), "| env:", {k: v for k, v in os.environ.items() if k.startswith("MGODPL_")})
