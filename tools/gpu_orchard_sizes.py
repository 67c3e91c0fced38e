"""Dev tool: the orchard motion batch at shard sizes (strong-scaling view) and split over 2 streams (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
orchard = mg.scenes.make_orchard(); mesh = orchard.collision_mesh(True)
scene = mg.Scene.from_mesh(mesh)
model = mg.robots.create_procedural_robot_model(1.0, (0,)); robot = model.to_device()
dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
mm = 1_000_000
gen = torch.Generator(device=dev).manual_seed(5)
u = torch.rand((2 * mm, 6), generator=gen, device=dev, dtype=torch.float64)
st = torch.zeros((2 * mm, 9), dtype=torch.float64, device=dev)
st[:, 0], st[:, 1], st[:, 2] = (u[:, 0] * 2 - 1) * 10.0, (u[:, 1] * 2 - 1) * 5.0, u[:, 2] * 10.0
yaw = u[:, 3] * (2 * np.pi); st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
st[:, 7], st[:, 8] = (u[:, 4] * 2 - 1) * (np.pi / 2), (u[:, 5] * 2 - 1) * (np.pi / 2)
a, b = st[:mm].contiguous(), st[mm:].contiguous()
d_h = torch.zeros((mm + 31) // 32, dtype=torch.int32, device=dev); d_t = torch.zeros(mm, dtype=torch.float64, device=dev)
d_s = torch.zeros(mm, dtype=torch.int32, device=dev)
side = torch.cuda.Stream(device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in (125_000, 250_000, 500_000, 1_000_000):
    def one():
        mg.capi.check_motions_device(scene, robot, a, b, n, 9, 2, d_h, 0.1, d_t, d_s, stream=stream)

    def two():
        h = (n // 2 + 31) // 32 * 32
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            mg.capi.check_motions_device(scene, robot, a[h:], b[h:], n - h, 9, 2, d_h[h // 32:], 0.1, d_t[h:], d_s[h:], stream=side.cuda_stream)
        mg.capi.check_motions_device(scene, robot, a, b, h, 9, 2, d_h, 0.1, d_t, d_s, stream=stream)
        torch.cuda.current_stream().wait_stream(side)
    t1, t2 = timed(one), timed(two)
    print(f"orchard motions n={n}: one stream {t1:.3f} ms ({n / t1 / 1e3:.1f} M/s), two streams {t2:.3f} ms; ideal share of 1M: {n / mm:.3f}", flush=True)
