"""Dev tool: config 3 only (500 fruit x 1000 goal samples), for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE); mesh = tree.collision_mesh(len(sys.argv) < 2)
scene = mg.Scene.from_mesh(mesh, counters=True)
model = mg.robots.create_procedural_robot_model(1.0, (0,)); robot = model.to_device()
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
f, s = len(tree.fruit_centers), 1000
d_t = torch.from_numpy(tree.fruit_centers.copy()).to(dev)
d_gh = torch.zeros(f * ((s + 31) // 32), dtype=torch.int32, device=dev); d_cnt = torch.zeros(f, dtype=torch.int32, device=dev)
for _ in range(3):
    mg.capi.sample_goals_device(scene, robot, d_t, f, s, d_gh, d_collisions=d_cnt, stream=stream)
torch.cuda.synchronize(); scene.counters(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    mg.capi.sample_goals_device(scene, robot, d_t, f, s, d_gh, d_collisions=d_cnt, stream=stream)
e1.record(); torch.cuda.synchronize()
c = scene.counters()
print("goals 500x1000: %.3f ms, collision rate %.4f" % (e0.elapsed_time(e1) / 5, d_cnt.cpu().numpy().sum() / (f * s)), {k: v / 5 for k, v in c.items()})
