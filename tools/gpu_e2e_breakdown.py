"""Dev tool: where does the host-buffer (e2e) call spend its time?"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgodpl_b200 as mg
import bench
tree = mg.scenes.make_tree(**bench.TREE); mesh = tree.collision_mesh(True)
scene = mg.Scene.from_mesh(mesh)
model, a, b = bench.make_inputs(mg, 0); robot = model.to_device()
m = len(a)
pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
na, nb = pa.numpy(), pb.numpy()
words = np.zeros((m + 31)//32, np.uint32); toi = np.zeros(m); steps = np.zeros(m, np.uint32)
L = mg.capi.lib()
def raw():
    rc = L.mgodpl_check_motions(scene.h, robot.h, na.ctypes.data_as(C.POINTER(C.c_double)), nb.ctypes.data_as(C.POINTER(C.c_double)),
        C.c_size_t(m), C.c_size_t(9), 2, C.c_double(0.1), words.ctypes.data_as(C.POINTER(C.c_uint32)), toi.ctypes.data_as(C.POINTER(C.c_double)),
        steps.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_void_p(0))
    assert rc == 0
for _ in range(3): raw()
t0 = time.perf_counter()
for _ in range(20): raw()
print("raw C ABI call (pinned in, pageable out): %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
for _ in range(3): mg.capi.check_motions(scene, robot, na, nb)
t0 = time.perf_counter()
for _ in range(20): mg.capi.check_motions(scene, robot, na, nb)
print("python wrapper call: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
d = torch.empty_like(pa, device="cuda")
for _ in range(3): d.copy_(pa, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): d.copy_(pa, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("H2D 7.2 MB pinned: %.3f ms (%.1f GB/s)" % (dt * 1e3, a.nbytes / dt / 1e9))
pag = a.copy(); tp = torch.from_numpy(pag)
t0 = time.perf_counter()
for _ in range(20): d.copy_(tp)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("H2D 7.2 MB pageable: %.3f ms (%.1f GB/s)" % (dt * 1e3, a.nbytes / dt / 1e9))
