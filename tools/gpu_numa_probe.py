"""Dev tool: does the host-buffer call depend on which NUMA node the calling thread and its pinned buffers sit on?
Prints the host topology as this process sees it, then times the 14.4 MB input copy and the host-buffer motion call with the
process bound to each NUMA node's CPUs in turn (pinned buffers allocated after binding, so first touch places them there)."""
import ctypes
import glob
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import mgodpl_b200 as mg


def cpulist(text):
    out = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


def ranges(cpus):
    cpus, out, i = sorted(cpus), [], 0
    while i < len(cpus):
        j = i
        while j + 1 < len(cpus) and cpus[j + 1] == cpus[j] + 1:
            j += 1
        out.append(f"{cpus[i]}-{cpus[j]}" if j > i else f"{cpus[i]}")
        i = j + 1
    return ",".join(out)


libc = ctypes.CDLL(None, use_errno=True)
allowed = os.sched_getaffinity(0)
print("allowed cpus:", len(allowed), ranges(allowed), "| running on cpu", libc.sched_getcpu())
nodes = {}
for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    try:
        nodes[int(d.rsplit("node", 1)[1])] = cpulist(open(d + "/cpulist").read())
    except Exception as e:
        print("node read failed", d, e)
for n, c in nodes.items():
    print(f"node {n}: {len(c)} cpus {ranges(c)} | allowed here: {len(c & allowed)}")
props = torch.cuda.get_device_properties(0)
bus = None
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    bus = bus.lower()
    if len(bus.split(":")[0]) == 8:
        bus = bus[4:]
    print("gpu 0:", props.name, "pci", bus)
    for f in ("numa_node", "local_cpulist", "current_link_speed", "current_link_width"):
        try:
            print(f"  {f}:", open(f"/sys/bus/pci/devices/{bus}/{f}").read().strip())
        except Exception as e:
            print(f"  {f}: unreadable ({e})")
except Exception as e:
    print("nvml / sysfs:", e)
try:
    print("numa_balancing:", open("/proc/sys/kernel/numa_balancing").read().strip())
except Exception:
    pass

tree = mg.scenes.make_tree(**bench.TREE)
scene = mg.Scene.from_mesh(tree.collision_mesh(True))
model, a, b = bench.make_inputs(mg, 0)
robot = model.to_device()
m = len(a)
res = mg.capi.motion_result_buffers(m)


def measure(tag):
    pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    na, nb = pa.numpy(), pb.numpy()
    d = torch.empty_like(pa, device="cuda")
    cp = []
    for _ in range(30):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.copy_(pa, non_blocking=True)
        torch.cuda.synchronize()
        cp.append((time.perf_counter() - t0) * 1e3)
    ts = []
    for _ in range(260):
        t0 = time.perf_counter()
        mg.capi.check_motions(scene, robot, na, nb, 0.1, True, out=res)
        ts.append((time.perf_counter() - t0) * 1e3)
    cp2 = []
    for _ in range(30):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.copy_(pa, non_blocking=True)
        torch.cuda.synchronize()
        cp2.append((time.perf_counter() - t0) * 1e3)
    gbs = pa.numel() * 8 / (np.median(cp2) * 1e-3) / 1e9
    print(f"[{tag}] cpu {libc.sched_getcpu()} | 7.2 MB H2D median {np.median(cp):.3f} -> {np.median(cp2):.3f} ms ({gbs:.1f} GB/s) | "
          f"e2e first10 {np.median(ts[:10]):.3f}, calls 100-130 {np.median(ts[100:130]):.3f}, last30 {np.median(ts[-30:]):.3f} (min {min(ts):.3f}) ms",
          flush=True)


measure("as started")
for n, c in nodes.items():
    use = c & allowed
    if not use:
        print(f"[node {n}] no allowed cpus")
        continue
    os.sched_setaffinity(0, use)
    time.sleep(0.01)
    measure(f"bound to node {n}")
    # one cpu of the node only (the calling thread cannot migrate at all)
    os.sched_setaffinity(0, {min(use)})
    measure(f"bound to cpu {min(use)}")
os.sched_setaffinity(0, allowed)
measure("unbound again")
