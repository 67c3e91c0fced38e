"""Dev tool: A/B of the clearance grid on the bench workloads (run on the GPU box): device counters + timings."""
import sys, os
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
n = 10_000_000
gen = torch.Generator(device=dev).manual_seed(1)
u = torch.rand((n, 6), generator=gen, device=dev, dtype=torch.float64)
st = torch.zeros((n, 9), dtype=torch.float64, device=dev)
st[:, 0], st[:, 1], st[:, 2] = (u[:, 0] * 2 - 1) * 5.0, (u[:, 1] * 2 - 1) * 5.0, u[:, 2] * 10.0
yaw = u[:, 3] * (2 * np.pi)
st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
st[:, 7], st[:, 8] = (u[:, 4] * 2 - 1) * (np.pi / 2), (u[:, 5] * 2 - 1) * (np.pi / 2)
d_sh = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
ref = {}
for grid in (False, True):
    for counters in (True, False):
        scene = mg.Scene.from_mesh(mesh, counters=counters, clearance_grid=grid)
        info = scene.info()

        def motions():
            mg.capi.check_motions_device(scene, robot, d_a, d_b, m, 9, 2, d_hit, 0.1, d_toi, d_steps, stream=stream)

        def states():
            mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_sh, stream=stream)
        for name, fn, buf in (("motions100k", motions, d_hit), ("states10M", states, d_sh)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            if counters:
                scene.counters(reset=True)
            ts = []
            for k in range(10):
                flush.fill_(k)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            words = buf.cpu().numpy().copy()
            key = name
            if key in ref:
                assert np.array_equal(ref[key], words), f"{name}: grid={grid} changes the result"
            else:
                ref[key] = words
            msg = f"grid={int(grid)} {name}: ms={np.mean(ts):.4f} (min {np.min(ts):.4f})"
            if counters:
                c = {x: v / 10 for x, v in scene.counters().items()}
                msg += (f" [with counters] states={c['states_checked']:.4g} boxes={c['boxes_traversed']:.4g} records={c['nodes_visited']:.4g} "
                        f"rec/box={c['nodes_visited'] / max(c['boxes_traversed'], 1):.2f} exact={c['leaf_tests']:.4g}")
            print(msg, flush=True)
        if counters:
            print(f"   scene: build {info.build_ms:.2f} ms + grid {info.clearance_build_ms:.2f} ms, {info.clearance_cells} cells of {info.clearance_cell_size * 100:.2f} cm")
