"""Dev tool: one 125 k-motion / 1.25 M-state shard of the orchard batch (what a rank of 8 runs), for launch lists (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
orchard = mg.scenes.make_orchard(); mesh = orchard.collision_mesh(True)
scene = mg.Scene.from_mesh(mesh)
model = mg.robots.create_procedural_robot_model(1.0, (0,)); robot = model.to_device()
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
mm = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
ns = 10 * mm
gen = torch.Generator(device=dev).manual_seed(5)
u = torch.rand((ns, 6), generator=gen, device=dev, dtype=torch.float64)
st = torch.zeros((ns, 9), dtype=torch.float64, device=dev)
st[:, 0], st[:, 1], st[:, 2] = (u[:, 0] * 2 - 1) * 10.0, (u[:, 1] * 2 - 1) * 5.0, u[:, 2] * 10.0
yaw = u[:, 3] * (2 * np.pi); st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
st[:, 7], st[:, 8] = (u[:, 4] * 2 - 1) * (np.pi / 2), (u[:, 5] * 2 - 1) * (np.pi / 2)
a, b = st[:mm].contiguous(), st[mm:2 * mm].contiguous()
d_h = torch.zeros((mm + 31) // 32, dtype=torch.int32, device=dev); d_t = torch.zeros(mm, dtype=torch.float64, device=dev)
d_s = torch.zeros(mm, dtype=torch.int32, device=dev); d_sh = torch.zeros((ns + 31) // 32, dtype=torch.int32, device=dev)
for k in range(4):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    mg.capi.check_motions_device(scene, robot, a, b, mm, 9, 2, d_h, 0.1, d_t, d_s, stream=stream)
    e1.record()
    mg.capi.check_states_device(scene, robot, st, ns, 9, 2, d_sh, stream=stream)
    e2.record()
    torch.cuda.synchronize()
    print(f"motions {mm}: {e0.elapsed_time(e1):.4f} ms   states {ns}: {e1.elapsed_time(e2):.4f} ms", flush=True)
