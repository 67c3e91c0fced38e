#!/usr/bin/env python
"""Benchmark of the collision-query hot path (BASELINE.json metric: collision-checked states/s and motions/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over one batch of synthetic input:
  workload  BASELINE.json configs[1]: 100 000 straight-line motions between PRM-sampler states
            (generateUniformRandomState, h=5, v=10; src/planning/tsp_over_prm.cpp:569-571) validated with
            check_motion_collides at the reference's 0.1 step, early exit, toi reported, against the 200 k-triangle
            synthetic tree (trunk + leaves), default 0.75 m one-joint arm.
  value     collision-checked robot states per second, counted the way the reference would count them
            (a motion that first collides at step k costs k+1 state checks, a free one n_steps+1), inputs resident in
            HBM, device time from CUDA events on the launching stream, L2 flushed before every timed step.
  e2e       the same metric through the host-buffer C ABI call (pinned host inputs, H2D + kernel + D2H inside): the sum of the
            synchronous calls' wall times; clocks are queried between the calls, from the calling thread.
With N > 1 (torchrun) every rank holds a replica of the scene and validates its own 100 000 motions (weak scaling);
the packed result words are all-gathered over NCCL inside the step.

--impl reference times the reference's own CPU implementation of the same path (oracle/_ref: the reference's sources
compiled unmodified over the FCL/Eigen shim; falls back to the oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_MOTIONS = 100_000
ARM, JOINTS = 0.75, (0,)
TREE = dict(seed=42, trunk_triangles=50_000, leaf_triangles=150_000, n_fruit=500)
MAX_STEP = 0.1
GATHER_EVERY = int(os.environ.get("MGODPL_BENCH_GATHER_EVERY", "4"))  # N > 1: result words of this many consecutive batches share one all-gather
# our kernels inside the device-timed region, per step
LAUNCHES_PER_STEP = 7
KERNEL_NOTE = ("check_motions pipeline: k_motion_ranges + k_expand_overflow_ranges (idle unless a work list ran full) + 2 x (k_expand_motion_steps + k_traverse) + k_finish_motions; k_traverse is the "
               "dominant kernel (see profiles/)")
B_MOTION = 2 * 8 * (7 + 2) + 1 / 8 + 8      # two fp64 states in, 1 result bit + toi out = 152.125 B per motion
B_SCENE_PER_TRIANGLE = 48 + 64              # 3 x float4 vertices + ~2 nodes x 32 B per triangle
WORKLOAD = ("configs[1]: 100k straight-line motions/GPU, check_motion_collides @ step 0.1, early exit + toi, "
            "200k-triangle synthetic tree (trunk+leaves), 0.75 m 1-joint arm, PRM sampler h=5 v=10")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def make_inputs(mg, rank: int):
    model = mg.robots.create_procedural_robot_model(ARM, JOINTS)
    rng = mg.robots.Mt19937(42 + rank)
    s = mg.robots.generate_uniform_random_states(model, rng, 2 * M_MOTIONS, 5.0, 10.0)
    return model, s[:M_MOTIONS].copy(), s[M_MOTIONS:].copy()


def ref_equivalent_states(hit, toi, n_steps):
    """State checks the reference's sequential loop performs (collision_detection.cpp:94-103)."""
    n = np.maximum(n_steps, 1)
    first = np.rint(np.nan_to_num(toi) * n).astype(np.int64)
    return np.where(hit, np.where(n_steps > 0, first + 1, 1), n_steps + 1)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING warm-up + timed region (B200_PROFILING.md recipe).  NVML is polled
    every few ms (the timed region lasts tens of ms); nvidia-smi is the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.sm, self.max_sm, self.mask, self.stop_flag, self.how = index, [], None, 0, threading.Event(), "nvml"
        self.ready = threading.Event()  # set after the first sample: the caller starts its warm-up only then
        self.period = 0.0005            # the timed region of a 20-step run lasts a few ms
        self.paused = threading.Event() # set: the thread stops polling; the owner samples with sample_now() instead
        self.lock = threading.Lock()    # held around a query, so that pause() returns only when none is in flight
        self._query = None

    def mark(self):
        """Forget what was sampled so far (called right before the warm-up steps)."""
        self.ready.wait(20.0)
        self.sm, self.mask = [], 0

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self._query = lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), int(reasons(h)))
            while not self.stop_flag.is_set():
                if not self.paused.is_set():
                    with self.lock:
                        sm, mask = self._query()
                        self.sm.append(sm)
                        self.mask |= mask
                self.ready.set()
                self.stop_flag.wait(self.period)
            return
        except Exception:
            self.how = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag.is_set():
            if self.paused.is_set():
                self.stop_flag.wait(0.05)
                continue
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")]
                if r[0].isdigit():
                    self.sm.append(int(r[0]))
                    self.max_sm = int(r[1]) if r[1].isdigit() else self.max_sm
                    for i, bit in enumerate(bits):
                        if r[2 + i].lower().startswith("active"):
                            self.mask |= bit
            except Exception:
                pass
            self.ready.set()
            self.stop_flag.wait(0.05)

    def pause(self):
        """No more polling from the thread (NVML queries take driver locks that CUDA calls of other threads wait for: polled
        every 2 ms they made the host-buffer calls 15 % slower, every 0.5 ms 2.5x).  Returns once no query is in flight."""
        self.paused.set()
        with self.lock:
            self.sm, self.mask = [], 0

    def sample_now(self):
        """One sample from the calling thread (between two host-buffer calls: it cannot hold up a CUDA call of this thread)."""
        if self._query is None:
            return
        try:
            sm, mask = self._query()
            self.sm.append(sm)
            self.mask |= mask
        except Exception:
            pass

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["unsampled"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": [n for b, n in self.REASONS.items() if self.mask & b],
                "samples": len(sm), "sm_mhz_min": sm[0], "how": self.how}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import mgodpl_b200 as mg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream

    sampler = ClockSampler(local)  # started early: NVML initialisation takes longer than a short timed region
    sampler.start()
    tree = mg.scenes.make_tree(**TREE)
    mesh = tree.collision_mesh(True)
    scene = mg.Scene.from_mesh(mesh)
    info = scene.info()
    model, a, b = make_inputs(mg, rank)
    robot = model.to_device()
    stride, js, m = a.shape[1], a.shape[1] - 7, len(a)
    words = (m + 31) // 32

    # device-resident inputs / outputs
    d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    d_hit = torch.zeros(words, dtype=torch.int32, device=dev)
    d_toi = torch.zeros(m, dtype=torch.float64, device=dev)
    d_steps = torch.zeros(m, dtype=torch.int32, device=dev)
    d_unc = torch.zeros(words, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)  # 256 MiB > 126 MB of L2
    # N > 1: the only exchange on this path is the gather of the packed result words (12.5 KB per rank and batch).  It is
    # latency-bound and nothing waits for it inside a batch, so the words of GATHER_EVERY consecutive batches travel in one
    # asynchronous all-gather (NCCL's own stream), overlapping the following batches' kernels; two such groups of result
    # buffers alternate so that a gather never reads words a later step is rewriting.
    G = GATHER_EVERY if world > 1 else 1
    hit_groups = [torch.zeros((G, words), dtype=torch.int32, device=dev) for _ in range(2)]
    gathered = [torch.empty((world, G, words), dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else None
    pending = [None, None]
    d_hit = hit_groups[0][0]

    def step(k=0):
        group, slot = (k // G) & 1, k % G
        if world > 1 and slot == 0 and pending[group] is not None:
            pending[group].wait()  # the gather that last read this group of buffers (two groups ago)
            pending[group] = None
        mg.capi.check_motions_device(scene, robot, d_a, d_b, m, stride, js, hit_groups[group][slot], MAX_STEP, d_toi, d_steps, d_unc,
                                     stream=stream)
        if world > 1 and slot == G - 1:
            pending[group] = dist.all_gather_into_tensor(gathered[group], hit_groups[group], async_op=True)

    def drain(last_step=None):
        if world > 1 and last_step is not None and (last_step + 1) % G:  # a partly filled group at the end still has to travel
            group = (last_step // G) & 1
            if pending[group] is not None:
                pending[group].wait()
            pending[group] = dist.all_gather_into_tensor(gathered[group], hit_groups[group], async_op=True)
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    graph_note = "off (--no-graph)"
    if args.graph:
        # The device-buffer call never blocks the host or allocates once its scratch exists, so it can be captured: the step's
        # launches are replayed from CUDA graphs (one per result buffer), which takes the host out of a 0.3 ms step.  The timed
        # work is the same C-ABI call, recorded once.  Falls back to plain launches if the capture fails.
        plain_step = step
        try:
            step(0)
            drain(0)
            torch.cuda.synchronize()
            graphs = {}
            cap = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(cap):  # the library keeps one scratch area per stream: let it allocate before the capture starts
                mg.capi.check_motions_device(scene, robot, d_a, d_b, m, stride, js, hit_groups[0][0], MAX_STEP, d_toi, d_steps, d_unc, stream=cap.cuda_stream)
            torch.cuda.synchronize()
            for key in [(g_, s_) for g_ in range(2) for s_ in range(G)]:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                    mg.capi.check_motions_device(scene, robot, d_a, d_b, m, stride, js, hit_groups[key[0]][key[1]], MAX_STEP, d_toi, d_steps, d_unc,
                                                 stream=torch.cuda.current_stream().cuda_stream)
                graphs[key] = g
            torch.cuda.synchronize()

            def step(k=0):  # noqa: F811
                group, slot = (k // G) & 1, k % G
                if world > 1 and slot == 0 and pending[group] is not None:
                    pending[group].wait()
                    pending[group] = None
                graphs[(group, slot)].replay()
                if world > 1 and slot == G - 1:
                    pending[group] = dist.all_gather_into_tensor(gathered[group], hit_groups[group], async_op=True)
            graph_note = "the step's C-ABI call captured once per result buffer and replayed (CUDA graph)"
        except Exception as e:  # keep measuring with plain launches
            torch.cuda.synchronize()
            step = plain_step
            graph_note = f"capture failed, plain launches: {e!r}"[:200]
    sampler.mark()
    for k in range(args.warmup):
        step(k)
    drain(args.warmup - 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_host = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k)  # evict inputs and scene from L2 (outside the timed events)
        ev[k][0].record()
        step(k)
        ev[k][1].record()
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps  # how long the host needs to enqueue one step (it runs ahead of the GPU)
    drain(args.steps - 1)  # the stream now waits for the gathers still in flight ...
    tail = torch.cuda.Event(enable_timing=True)
    tail.record()  # ... and this event closes the timed region after them
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks_timed = sampler.summary()  # warm-up + timed region
    sampler.pause()                   # the e2e calls below are sampled from this thread, between calls (see ClockSampler.pause)
    d_hit = hit_groups[((args.steps - 1) // G) & 1][(args.steps - 1) % G]
    step_ms = [s.elapsed_time(e) for s, e in ev]
    total_ms = torch.tensor([sum(step_ms) + ev[-1][1].elapsed_time(tail)], dtype=torch.float64, device=dev)
    by_rank, clocks_by_rank = None, None
    if world > 1:
        # per rank: [whole timed region / steps, mean of the per-step events (what the rank's own kernels took)]
        mine = torch.tensor([float(total_ms.item()) / args.steps, float(np.mean(step_ms))], dtype=torch.float64, device=dev)
        all_ms = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(all_ms, mine)
        by_rank = [[round(float(x[0].item()), 4), round(float(x[1].item()), 4)] for x in all_ms]
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        cs = clocks_timed
        mine = torch.tensor([float(cs.get("sm_mhz") or 0), float(cs.get("sm_mhz_min") or 0)], dtype=torch.float64, device=dev)
        all_clk = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(all_clk, mine)
        clocks_by_rank = [[int(x[0].item()), int(x[1].item())] for x in all_clk]
    total_ms = float(total_ms.item())

    hit = mg.capi.unpack_bits(d_hit.cpu().numpy().view(np.uint32), m)
    toi, n_steps = d_toi.cpu().numpy(), d_steps.cpu().numpy().astype(np.int64)
    states_per_batch = torch.tensor([float(ref_equivalent_states(hit, toi, n_steps).sum())], dtype=torch.float64, device=dev)
    motions_total = torch.tensor([float(m)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(states_per_batch)
        dist.all_reduce(motions_total)
    states_all, motions_all = float(states_per_batch.item()), float(motions_total.item())
    value = states_all * args.steps / (total_ms * 1e-3)

    # ---- e2e: host-buffer C ABI call, pinned inputs, H2D + kernel + D2H in the timed region -------------------------
    pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    na, nb = pa.numpy(), pb.numpy()
    res = mg.capi.motion_result_buffers(m)  # caller-owned result arrays, reused across calls
    # Warm-up of the host-buffer path: at least 100 calls, then on until the call time has settled (the last eight within 3 % of
    # each other) or 400 calls / 0.8 s have passed.  On some hosts the first 50-100 calls after a process starts using the host
    # link take up to 2.5x as long and speed up gradually (seen as 1.5 -> 0.6 ms over 70 ms with SM clocks at maximum
    # throughout; tools/gpu_e2e_trace.py): the timed calls measure the steady state, and every one of them is reported.
    e2e_warm, t_warm = [], time.perf_counter()
    while len(e2e_warm) < 400 and (len(e2e_warm) < 3 or time.perf_counter() - t_warm < 0.8):
        t1 = time.perf_counter()
        mg.capi.check_motions(scene, robot, na, nb, MAX_STEP, True, stream=stream, out=res)
        e2e_warm.append((time.perf_counter() - t1) * 1e3)
        if len(e2e_warm) >= 100 and max(e2e_warm[-8:]) <= 1.03 * min(e2e_warm[-8:]):
            break
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 10))
    e2e_calls = []
    sampler.sm, sampler.mask = [], 0
    sampler.sample_now()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        mg.capi.check_motions(scene, robot, na, nb, MAX_STEP, True, stream=stream, out=res)  # returns with the results in host memory
        e2e_calls.append((time.perf_counter() - t1) * 1e3)
        sampler.sample_now()  # clocks + throttle reasons right after the call, outside its timer
    torch.cuda.synchronize()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    # every call is synchronous (inputs from pinned host memory in, results in host memory when it returns): the e2e time is the
    # sum of the calls' own wall times; the NVML queries between them are not part of any call
    e2e_s = torch.tensor([sum(e2e_calls) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    assert np.array_equal(mg.capi.unpack_bits(res["words"], m), hit), "host-buffer and device-buffer paths disagree"
    assert np.array_equal(res["n_steps"].astype(np.int64), n_steps) and np.array_equal(res["toi"], toi, equal_nan=True)
    clocks_e2e = sampler.summary()
    sampler.stop_flag.set()
    sampler.join()
    e2e_value = states_all * e2e_steps / float(e2e_s.item())

    strong = None
    if not args.no_extra:
        try:
            strong = orchard_strong_scaling(mg, torch, dist, dev, stream, rank, world)
        except Exception as e:  # never at the cost of the headline line
            strong = {"error": repr(e)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    ms_per_step = total_ms / args.steps
    # SURVEY §8d accounting: 152.125 B per motion + 112 B per triangle, once per launch
    algo_bytes = m * B_MOTION + B_SCENE_PER_TRIANGLE * int(info.n_triangles)
    achieved = algo_bytes / (np.mean(step_ms) * 1e-3) / 1e9
    # what this implementation actually has to touch once per launch: the same inputs/outputs + the step-count array, its own
    # traversal records and both triangle copies (mgodpl_scene_info.traversal_bytes) and the clearance grid
    own_bytes = m * (B_MOTION + 4) + int(info.traversal_bytes) + int(info.clearance_cells)
    traffic, traffic_src, traffic_kept = None, None, {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            traffic, traffic_src = tj.get("k_check_motions_dram_bytes_per_launch"), tj.get("source")
            traffic_kept = tj.get("l2_kept_between_kernels", {})
    except Exception:
        pass
    line = {
        "metric": "collision_checked_states_per_sec", "value": value, "unit": "states/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "motions_per_gpu": m, "triangles": int(info.n_triangles), "bvh_nodes": int(info.n_nodes),
                   "l2": "flushed before every timed step (256 MiB write)", "parallelism": f"replicated scene x{world}, batch sharded" + (f", result words all-gathered every {G} batches (async)" if world > 1 else "")},
        "ms_per_step_by_rank": by_rank, "host_enqueue_ms_per_step": host_enqueue_ms, "launch_mode": graph_note,
        "motions_per_sec": motions_all * args.steps / (total_ms * 1e-3),
        "states_checked_per_motion": states_all / motions_all, "collision_rate": float(hit.mean()),
        "e2e": {"value": e2e_value, "unit": "states/s", "h2d_bytes_per_step": int(2 * m * stride * 8),
                "d2h_bytes_per_step": int(2 * words * 4 + m * 4 + m * 8), "steps": e2e_steps,
                "ms_per_call": [round(x, 4) for x in e2e_calls], "timed": "sum of the synchronous calls' wall times, max over ranks",
                "wall_ms_with_clock_queries_between_calls": round(e2e_wall_ms, 4), "warmup_calls": len(e2e_warm),
                "warmup_first_last_ms": [round(e2e_warm[0], 4), round(e2e_warm[-1], 4)]},
        # our kernels inside the device-timed region, per step: k_motion_ranges, 2 x (k_expand_motion_steps, k_traverse), k_finish_motions
        "gpu_launches": LAUNCHES_PER_STEP * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
                     # the same step with L2 contents kept between its kernels (as in the timed run): what really comes from DRAM
                     "traffic_reads_l2_kept_between_kernels": traffic_kept.get("dram_bytes_read_per_step"),
                     "kernel": KERNEL_NOTE,
                     "algorithmic_bytes_per_launch": algo_bytes, "algorithmic_bytes": "152.125 B x motions + 112 B x triangles (SURVEY 8d)",
                     "own_accounting_bytes_per_launch": own_bytes, "own_accounting_frac": own_bytes / (np.mean(step_ms) * 1e-3) / 1e9 / peak},
        "clocks": dict(clocks_timed, by_rank_median_min=clocks_by_rank, during_e2e=clocks_e2e),
        "scene": {"build_ms": info.build_ms, "sah_cost": info.sah_cost, "max_depth": int(info.max_depth),
                  "device_bytes": int(info.device_bytes)},
    }
    # Secondary figures: device-side work counters of one untimed step on a counters-enabled copy of the
    # scene — what actually limits this path is L1/L2-level traversal traffic and FP32 issue, not HBM.
    try:
        cscene = mg.Scene.from_mesh(mesh, counters=True)
        mg.capi.check_motions_device(cscene, robot, d_a, d_b, m, stride, js, hit_groups[0][0], MAX_STEP, d_toi, d_steps, d_unc, stream=stream)
        torch.cuda.synchronize()
        c = cscene.counters()
        rec_bytes = 128 if int(info.n_wide_nodes) else 64
        line["secondary"] = {
            "states_expanded_on_device": c["states_checked"], "boxes_traversed": c["boxes_traversed"],
            "records_visited": c["nodes_visited"], "exact_fp64_leaf_tests": c["leaf_tests"],
            "max_records_one_box": c["max_nodes_per_box"],
            "reference_equivalent_states": states_all / world,
            "traversal_bytes_per_reference_state": rec_bytes * c["nodes_visited"] / (states_all / world),
            "traversal_bytes_per_box": rec_bytes * c["nodes_visited"] / max(c["boxes_traversed"], 1),
            "note": "L1/L2-level bytes (record size x records visited); ncu L2 hit rate, FMA-pipe and issue utilisation per kernel "
                    "are in profiles/r02_final_ncu_full_motion_step.txt"}
        del cscene
    except Exception as e:  # never at the cost of the headline line
        line["secondary"] = {"error": repr(e)}
    if not args.no_cpu_baseline:
        if world == 1:  # the CPU baseline is timed on rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_baseline(mesh, a, b, use_ref=False)
            # one host thread: how a single reference call executes (collision_detection.cpp:57-105 is sequential)
            line["cpu_baseline_1thread"] = cpu_baseline(mesh, a, b, use_ref=False, budget_s=6.0, cores=1)
        try:
            line["parity"] = parity_sample(mesh, a, b, hit)
        except Exception as e:
            line["parity"] = {"error": repr(e)}
    if strong is not None:
        line["strong_scaling"] = strong
    if not args.no_extra:
        line["extra"] = extra_workloads(mg, torch, scene, tree, dev, stream, cpu_leg=not args.no_cpu_baseline and world == 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


ORCHARD_STATES, ORCHARD_MOTIONS = 10_000_000, 1_000_000


def orchard_strong_scaling(mg, torch, dist, dev, stream, rank, world, steps=6):
    """BASELINE configs[4]: the 8-tree orchard row (~2 M triangles, makeSingleRowOrchard, TreeMeshes.cpp:197-206) replicated
    on every GPU; ONE fixed batch of 10 M state checks + 1 M motions split over the ranks (strong scaling).  States: equal
    32-aligned shards.  Motions: shards balanced by the predicted state checks per motion, ceil(d / 0.1) + 1, which is known
    before launch (collision_detection.cpp:88-92).  One all-gather per batch brings every rank's packed result words and
    toi values to all ranks.  Timed on the device (CUDA events), max over ranks.  Rank 0 afterwards runs the whole batch
    alone for the in-run single-GPU reference."""
    orchard = mg.scenes.make_orchard()
    omesh = orchard.collision_mesh(True)
    oscene = mg.Scene.from_mesh(omesh)
    oinfo = oscene.info()
    model = mg.robots.create_procedural_robot_model(1.0, (0,))
    robot = model.to_device()
    n, mm = ORCHARD_STATES, ORCHARD_MOTIONS
    gen = torch.Generator(device=dev).manual_seed(5)  # same seed on every rank: the same global batch everywhere
    u = torch.rand((n, 6), generator=gen, device=dev, dtype=torch.float64)
    st = torch.zeros((n, 9), dtype=torch.float64, device=dev)
    st[:, 0], st[:, 1], st[:, 2] = (u[:, 0] * 2 - 1) * 10.0, (u[:, 1] * 2 - 1) * 5.0, u[:, 2] * 10.0   # SURVEY §8d config 5 box
    yaw = u[:, 3] * (2 * np.pi)
    st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
    st[:, 7], st[:, 8] = (u[:, 4] * 2 - 1) * (np.pi / 2), (u[:, 5] * 2 - 1) * (np.pi / 2)
    del u
    d_ma, d_mb = st[:mm].contiguous(), st[mm:2 * mm].contiguous()
    t0 = time.perf_counter()
    costs = mg.sharding.predicted_motion_costs(d_ma.cpu().numpy(), d_mb.cpu().numpy(), MAX_STEP)
    partition_ms = (time.perf_counter() - t0) * 1e3

    side = torch.cuda.Stream(device=dev)

    def plan(w):
        sb = [mg.sharding.shard_range(n, r, w) for r in range(w)]
        mb = mg.sharding.balanced_bounds(costs, w)
        pad_s = max(mg.sharding.words_for(hi - lo) for lo, hi in sb)
        pad_m = max(hi - lo for lo, hi in mb)
        pad_mw = (mg.sharding.words_for(pad_m) + 1) // 2 * 2
        pad_s = (pad_s + 1) // 2 * 2  # keep the fp64 toi block 8-byte aligned
        return sb, mb, pad_s, pad_m, pad_mw

    launch_modes = []

    def run(w, r, gather):
        """Time `steps` batches on shard r of w; returns (ms per batch, the gathered staging buffers of the last batch, plan, n_steps)."""
        sb, mb, pad_s, pad_m, pad_mw = plan(w)
        (slo, shi), (mlo, mhi) = sb[r], mb[r]
        # one staging buffer per rank: [state words | motion words | toi], gathered with ONE collective per batch.  Two sets
        # alternate, so the (asynchronous) gather of batch i overlaps the kernels of batch i + 1.
        words_total = pad_s + pad_mw + 2 * pad_m
        local = [torch.zeros(words_total, dtype=torch.int32, device=dev) for _ in range(2)]
        everyone = [torch.empty(w * words_total, dtype=torch.int32, device=dev) for _ in range(2)] if gather else None
        d_steps = torch.zeros(max(mhi - mlo, 1), dtype=torch.int32, device=dev)
        d_states, a_sh, b_sh = st[slo:shi], d_ma[mlo:mhi], d_mb[mlo:mhi]
        pending = [None, None]

        def kernels(buf):
            d_sw, d_mw, d_toi = buf[:pad_s], buf[pad_s:pad_s + pad_mw], buf[pad_s + pad_mw:].view(torch.float64)
            # the state half and the motion half of a batch are independent: they run on two streams (a small shard is bound by
            # the latency of its few launches, not by throughput), joined before the results travel
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            if shi > slo:
                with torch.cuda.stream(side):
                    mg.capi.check_states_device(oscene, robot, d_states, shi - slo, 9, 2, d_sw, stream=side.cuda_stream)
            if mhi > mlo:
                mg.capi.check_motions_device(oscene, robot, a_sh, b_sh, mhi - mlo, 9, 2, d_mw, MAX_STEP, d_toi, d_steps, stream=cur.cuda_stream)
            cur.wait_stream(side)

        # both library calls of a batch (neither blocks the host nor allocates once its scratch exists) are captured into one CUDA
        # graph per staging buffer and replayed; plain launches if the capture fails (e.g. a batch so large that it needs the
        # survivor-count read-back)
        graphs = [None, None]
        try:
            # (a shard whose worst case exceeds the library's 2^25-item box queue reads a survivor count back mid-call: not capturable)
            if ((mhi - mlo) * 32 + 4096) * 4 > 2 ** 25 or (shi - slo) * 4 > 2 ** 25:
                raise RuntimeError("shard too large to capture")
            cap = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(cap):
                kernels(local[0])
            torch.cuda.synchronize()
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                    kernels(local[i])
                graphs[i] = g
            torch.cuda.synchronize()
        except Exception:
            torch.cuda.synchronize()
            graphs = [None, None]

        def batch(k):
            buf = local[k & 1]
            if pending[k & 1] is not None:
                pending[k & 1].wait()  # the gather that last read this staging buffer (two batches ago)
                pending[k & 1] = None
            if graphs[k & 1] is not None:
                graphs[k & 1].replay()
            else:
                kernels(buf)
            if gather:
                pending[k & 1] = dist.all_gather_into_tensor(everyone[k & 1], buf, async_op=True)

        def drain():
            for i in range(2):
                if pending[i] is not None:
                    pending[i].wait()
                    pending[i] = None
        for k in range(2):
            batch(k)
        drain()
        torch.cuda.synchronize()
        if gather:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            batch(k)
        drain()  # the stream waits for the gathers still in flight: they are inside the timed region
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if gather:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        last = (steps - 1) & 1
        full = everyone[last].view(w, words_total) if gather else local[last].view(1, words_total)
        launch_modes.append("graph replay" if graphs[0] is not None else "plain launches")
        return float(ms.item()), full, (sb, mb, pad_s, pad_m, pad_mw), d_steps

    ms_n, full, (sb, mb, pad_s, pad_m, pad_mw), _ = run(world, rank, world > 1)
    # reassemble the global results from the gathered staging buffers (every rank holds all of them)
    s_words = torch.cat([full[r, :mg.sharding.words_for(hi - lo)] for r, (lo, hi) in enumerate(sb)])
    m_words = torch.cat([full[r, pad_s:pad_s + mg.sharding.words_for(hi - lo)] for r, (lo, hi) in enumerate(mb)])
    toi = torch.cat([full[r, pad_s + pad_mw:].view(torch.float64)[:hi - lo] for r, (lo, hi) in enumerate(mb)])
    s_hit = mg.capi.unpack_bits(s_words.cpu().numpy().view(np.uint32), n)
    m_hit = mg.capi.unpack_bits(m_words.cpu().numpy().view(np.uint32), mm)
    toi = toi.cpu().numpy()
    shard_costs = [float(costs[lo:hi].sum()) for lo, hi in mb]
    out = {"workload": "configs[4]: 8-tree orchard row, 10M state checks + 1M motions (sampler box h_x=10 h_y=5 v=10, 1.0 m 1-joint arm), "
                       "one fixed batch sharded over the GPUs, scene replicated",
           "scaling": "strong", "n_gpus": world, "steps": steps, "triangles": int(oinfo.n_triangles), "scene_bytes": int(oinfo.device_bytes),
           "scene_build_ms": oinfo.build_ms, "clearance_grid_build_ms": oinfo.clearance_build_ms,
           "launch_mode": launch_modes[0], "ms_per_batch": ms_n, "states_per_sec": n / (ms_n * 1e-3), "motions_per_sec": mm / (ms_n * 1e-3),
           "state_collision_rate": float(s_hit.mean()), "motion_collision_rate": float(m_hit.mean()),
           "gathered_bytes_per_rank": int(4 * (pad_s + pad_mw + 2 * pad_m)),
           "motion_shards": {"balanced_by": "predicted ceil(d/0.1)+1 per motion", "sizes": [hi - lo for lo, hi in mb],
                             "predicted_cost_max_over_mean": max(shard_costs) / (sum(shard_costs) / len(shard_costs)), "partition_ms_host": partition_ms}}
    if rank == 0:
        if world > 1:
            ms_1, full1, (sb1, mb1, ps1, pm1, pmw1), d_steps1 = run(1, 0, False)
            same = (np.array_equal(full1[0, :mg.sharding.words_for(n)].cpu().numpy(), s_words.cpu().numpy()) and
                    np.array_equal(full1[0, ps1:ps1 + mg.sharding.words_for(mm)].cpu().numpy(), m_words.cpu().numpy()) and
                    np.array_equal(full1[0, ps1 + pmw1:].view(torch.float64)[:mm].cpu().numpy(), toi, equal_nan=True))
            out["single_gpu_ms_per_batch_same_run"] = ms_1
            out["efficiency_vs_single_gpu"] = ms_1 / (world * ms_n)
            out["sharded_results_identical_to_single_gpu"] = bool(same)
        else:
            out["single_gpu_ms_per_batch_same_run"] = ms_n
            out["efficiency_vs_single_gpu"] = 1.0
        # the two halves on their own (rank 0, whole batch): states/s and reference-equivalent states/s of the motion half
        d_sw = torch.zeros(mg.sharding.words_for(n), dtype=torch.int32, device=dev)
        d_mw = torch.zeros(mg.sharding.words_for(mm), dtype=torch.int32, device=dev)
        d_mt = torch.zeros(mm, dtype=torch.float64, device=dev)
        d_ms = torch.zeros(mm, dtype=torch.int32, device=dev)

        def timed(fn, reps=3):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        t_s = timed(lambda: mg.capi.check_states_device(oscene, robot, st, n, 9, 2, d_sw, stream=stream))
        t_m = timed(lambda: mg.capi.check_motions_device(oscene, robot, d_ma, d_mb, mm, 9, 2, d_mw, MAX_STEP, d_mt, d_ms, stream=stream))
        sc = ref_equivalent_states(m_hit, toi, d_ms.cpu().numpy().astype(np.int64)).sum()
        out["single_gpu"] = {"states_10M": {"ms": t_s, "states_per_sec": n / (t_s * 1e-3), "hbm_frac": n * 72.125 / (t_s * 1e-3) / 1e9 / peaks()[0]},
                             "motions_1M": {"ms": t_m, "motions_per_sec": mm / (t_m * 1e-3), "states_per_sec": float(sc) / (t_m * 1e-3)}}
    del st, d_ma, d_mb
    return out


def extra_workloads(mg, torch, scene, tree, dev, stream, cpu_leg=True):
    """Secondary numbers (not the headline): 10 M state checks and 500 x 1000 goal samples on the same tree."""
    out = {}
    model = mg.robots.create_procedural_robot_model(1.0, (0,))
    robot = model.to_device()
    n = 10_000_000
    lo, hi = tree.trunk_mesh.aabb()
    h = max(hi[0] - lo[0], hi[1] - lo[1]) / 2 + 2.0
    gen = torch.Generator(device=dev).manual_seed(1)
    u = torch.rand((n, 6), generator=gen, device=dev, dtype=torch.float64)
    st = torch.zeros((n, 9), dtype=torch.float64, device=dev)
    st[:, 0] = (u[:, 0] * 2 - 1) * 5.0
    st[:, 1] = (u[:, 1] * 2 - 1) * 5.0
    st[:, 2] = u[:, 2] * 10.0
    yaw = u[:, 3] * (2 * np.pi)
    st[:, 5], st[:, 6] = torch.sin(yaw / 2), torch.cos(yaw / 2)
    st[:, 7] = (u[:, 4] * 2 - 1) * (np.pi / 2)
    st[:, 8] = (u[:, 5] * 2 - 1) * (np.pi / 2)
    d_hit = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
    for name, hh, vv in (("states_10M_prm_box_h5_v10", 5.0, 10.0), ("config1_states_10M_sampling_difficulty_box", h, float(lo[2] + 1.0))):
        st[:, 0] = (u[:, 0] * 2 - 1) * hh
        st[:, 1] = (u[:, 1] * 2 - 1) * hh
        st[:, 2] = u[:, 2] * vv
        for _ in range(2):
            mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_hit, stream=stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            mg.capi.check_states_device(scene, robot, st, n, 9, 2, d_hit, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[name] = {"states_per_sec": n / (ms * 1e-3), "ms": ms, "hbm_frac": n * 72.125 / (ms * 1e-3) / 1e9 / peaks()[0],
                     "collision_rate": float(mg.capi.unpack_bits(d_hit.cpu().numpy().view(np.uint32), n).mean())}
    # ---- the same 10 M states with a capsule arm and a convex-polytope end effector (north_star (3): link shapes beside boxes;
    #      the reference has no such shapes, so this line has no CPU leg) --------------------------------------------------------
    try:
        shaped = mg.robots.RobotModel()
        base = shaped.insert_link("flying_base", (0.8, 0.8, 0.2))
        arm = shaped.insert_link("stick")
        s45 = float(np.sin(np.pi / 4))
        shaped.add_capsule(arm, 0.035, 0.93, (0.0, 0.0, 0.0, s45, 0.0, 0.0, s45))  # axis along the link's y, like the reference's 1 m stick
        ee = shaped.insert_link("end_effector")
        shaped.add_convex_mesh(ee, [[-0.08, -0.1, -0.05], [0.08, -0.1, -0.05], [0.08, 0.1, -0.05], [-0.08, 0.1, -0.05], [0.0, 0.12, 0.09], [0.0, -0.06, 0.07]])
        ident = mg.capi.IDENTITY_TF
        shaped.insert_joint(base, arm, (0, 0.4, 0, 0, 0, 0, 1), (0, -0.5, 0, 0, 0, 0, 1), mg.capi.JOINT_REVOLUTE, (1.0, 0.0, 0.0), -np.pi / 2, np.pi / 2)
        shaped.insert_joint(arm, ee, (0, 0.5, 0, 0, 0, 0, 1), ident, mg.capi.JOINT_REVOLUTE, (0.0, 0.0, 1.0), -np.pi / 2, np.pi / 2)
        srobot = shaped.to_device()
        for _ in range(2):
            mg.capi.check_states_device(scene, srobot, st, n, 9, 2, d_hit, stream=stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            mg.capi.check_states_device(scene, srobot, st, n, 9, 2, d_hit, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out["states_10M_capsule_arm_convex_end_effector"] = {
            "states_per_sec": n / (ms * 1e-3), "ms": ms, "collision_rate": float(mg.capi.unpack_bits(d_hit.cpu().numpy().view(np.uint32), n).mean()),
            "note": "box base + capsule arm (r 0.035, l 0.93) + 6-vertex convex end effector, two revolute joints; same states as the line above"}
    except Exception as e:
        out["states_10M_capsule_arm_convex_end_effector"] = {"error": repr(e)}
    # ---- config 1 as BASELINE.json words it: ONE call with 10 000 states (host buffers in, booleans out) — launch-latency scale ----
    st10k = mg.robots.generate_uniform_random_states(model, mg.robots.Mt19937(42), 10_000, h, float(lo[2] + 1.0))
    for _ in range(3):
        mg.check_states(scene, robot, st10k, stream=stream)
    lat = []
    for _ in range(20):
        t0 = time.perf_counter()
        mg.check_states(scene, robot, st10k, stream=stream)
        lat.append((time.perf_counter() - t0) * 1e3)
    d10k = torch.from_numpy(st10k).to(dev)
    d10h = torch.zeros((10_000 + 31) // 32, dtype=torch.int32, device=dev)

    def timed(fn, reps=3, warm=2):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms10k = timed(lambda: mg.capi.check_states_device(scene, robot, d10k, 10_000, st10k.shape[1], st10k.shape[1] - 7, d10h, stream=stream), reps=20)
    out["config1_states_10k_single_call"] = {"host_call_ms_median": float(np.median(lat)), "host_call_ms_min": float(np.min(lat)),
                                             "device_ms": ms10k, "states_per_sec_host_call": 10_000 / (np.median(lat) * 1e-3),
                                             "note": "one mgodpl_check_states call, H2D + kernels + D2H + sync; bounded by launch latency, not throughput"}
    # ---- config 3: goal sampling, 500 fruit x 1000 samples, device-side draws; on the T+L scene (north_star extension) and on
    # the T scene the reference actually collides with (fcl_utils.cpp:58-60), the latter with the benchmark's robot sweep
    # {0.5, 0.75, 1.0} m x {H, V, HH, VV, HV} (goal_sample_success_rate.cpp:30-38) ---------------------------------------------
    f, s = len(tree.fruit_centers), 1000
    d_t = torch.from_numpy(tree.fruit_centers.copy()).to(dev)
    d_gh = torch.zeros(f * ((s + 31) // 32), dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(f, dtype=torch.int32, device=dev)
    ms = timed(lambda: mg.capi.sample_goals_device(scene, robot, d_t, f, s, d_gh, d_collisions=d_cnt, stream=stream))
    rate = float(d_cnt.cpu().numpy().sum() / (f * s))
    out["config3_goal_sampling_500x1000"] = {"scene": "T+L (trunk + leaves)", "samples_per_sec": f * s / (ms * 1e-3), "ms": ms,
                                             "collision_rate": rate, "valid_fraction": 1.0 - rate}
    try:
        tscene = mg.Scene.from_mesh(tree.collision_mesh(False))
        sweep, tot_ms = {}, 0.0
        for arm in (0.5, 0.75, 1.0):
            for name, jt in (("H", (0,)), ("V", (1,)), ("HH", (0, 0)), ("VV", (1, 1)), ("HV", (0, 1))):
                rb = mg.robots.create_procedural_robot_model(arm, jt).to_device()
                ms_r = timed(lambda: mg.capi.sample_goals_device(tscene, rb, d_t, f, s, d_gh, d_collisions=d_cnt, stream=stream))
                r = float(d_cnt.cpu().numpy().sum() / (f * s))
                sweep[f"{arm}m_{name}"] = {"ms": ms_r, "samples_per_sec": f * s / (ms_r * 1e-3), "valid_fraction": 1.0 - r}
                tot_ms += ms_r
        out["config3_goal_sampling_T_scene_robot_sweep"] = {"scene": "T (trunk only: the reference's collision set)", "fruit": f, "samples_per_fruit": s,
                                                            "robots": sweep, "samples_per_sec_all_robots": len(sweep) * f * s / (tot_ms * 1e-3)}
        del tscene
    except Exception as e:
        out["config3_goal_sampling_T_scene_robot_sweep"] = {"error": repr(e)}
    del st, u, d_hit
    # ---- config 4: PRM edge validation, 100k roadmap nodes, k = 10 (causal k-NN = the reference's incremental insertion) ----
    try:
        prm_model = mg.robots.create_procedural_robot_model(0.75, (0,))
        prm_robot = prm_model.to_device()
        cand = mg.robots.generate_uniform_random_states(prm_model, mg.robots.Mt19937(7), 104_000, 5.0, 10.0)
        free = ~mg.capi.check_states(scene, prm_robot, cand)
        nodes = np.ascontiguousarray(cand[free][:100_000])
        nn, k = len(nodes), 10
        d_nodes = torch.from_numpy(nodes).to(dev)
        d_idx = torch.zeros((nn, k), dtype=torch.int32, device=dev)
        d_nd = torch.zeros((nn, k), dtype=torch.float64, device=dev)
        ms_knn = timed(lambda: mg.capi.knn_device(d_nodes, nn, 9, 2, k, d_idx, d_nd, causal=True, stream=stream), reps=1, warm=1)
        idx = d_idx.cpu().numpy()
        src, col = np.nonzero(idx >= 0)
        d_ea = torch.from_numpy(np.ascontiguousarray(nodes[idx[src, col]])).to(dev)   # motion_collides(neighbour, state)
        d_eb = torch.from_numpy(np.ascontiguousarray(nodes[src])).to(dev)
        me = len(src)
        d_eh = torch.zeros((me + 31) // 32, dtype=torch.int32, device=dev)
        d_et = torch.zeros(me, dtype=torch.float64, device=dev)
        d_es = torch.zeros(me, dtype=torch.int32, device=dev)
        ms_e = timed(lambda: mg.capi.check_motions_device(scene, prm_robot, d_ea, d_eb, me, 9, 2, d_eh, MAX_STEP, d_et, d_es, stream=stream))
        eh = mg.capi.unpack_bits(d_eh.cpu().numpy().view(np.uint32), me)
        sc = ref_equivalent_states(eh, d_et.cpu().numpy(), d_es.cpu().numpy().astype(np.int64)).sum()
        out["config4_prm_edges_100k_nodes_k10"] = {"edges": me, "edges_per_sec": me / (ms_e * 1e-3), "states_per_sec": float(sc) / (ms_e * 1e-3),
                                                    "ms_edge_validation": ms_e, "ms_knn_100k_causal": ms_knn,
                                                    "edge_collision_rate": float(eh.mean()), "mean_steps": float(d_es.float().mean().item())}
        try:
            # ---- next step of the planner: shortest paths from every goal sample over the validated roadmap (runDijkstra per
            # source, tsp_over_prm.cpp:497-517), 1024 sources at once --------------------------------------------------------
            free_e = ~eh
            eu, ev = idx[src, col][free_e].astype(np.int64), src[free_e].astype(np.int64)
            ew = d_nd.cpu().numpy()[src, col][free_e]
            both_u, both_v, both_w = np.r_[eu, ev], np.r_[ev, eu], np.r_[ew, ew]
            order = np.argsort(both_u, kind="stable")
            row = np.zeros(nn + 1, np.int64)
            np.add.at(row, both_u + 1, 1)
            d_row = torch.from_numpy(np.cumsum(row).astype(np.uint32).view(np.int32)).to(dev)
            d_col = torch.from_numpy(both_v[order].astype(np.uint32).view(np.int32)).to(dev)
            d_w = torch.from_numpy(np.ascontiguousarray(both_w[order])).to(dev)
            g = 1024
            sources = np.random.default_rng(4).choice(nn, g, replace=False).astype(np.uint32)
            d_src = torch.from_numpy(sources.view(np.int32)).to(dev)
            d_dist = torch.empty((g, nn), dtype=torch.float64, device=dev)
            d_pred = torch.empty((g, nn), dtype=torch.int32, device=dev)
            sweeps = [0]

            def sssp():
                sweeps[0] = mg.capi.roadmap_shortest_paths_device(d_row, d_col, d_w, nn, d_src, g, d_dist, d_pred, stream=stream)
            ms_sp = timed(sssp, reps=3, warm=2)  # CUDA events around three calls (each synchronises for its convergence read-backs)
            sweeps = sweeps[0]
            sp = {"sources": g, "vertices": nn, "edges": int(free_e.sum()), "ms": ms_sp, "sources_per_sec": g / (ms_sp * 1e-3),
                  "labels_per_sec": g * nn / (ms_sp * 1e-3), "sweeps_enqueued": sweeps,
                  "reachable_fraction": float((d_dist[:8] < 1e300).float().mean().item())}
            if cpu_leg:
                from oracle import oracle as o  # CPU baseline leg only: the oracle's Dijkstra on the host cores, bounded sample
                threads = os.cpu_count() or 1
                n_cpu = min(g, 2 * threads)
                t0 = time.perf_counter()
                want, _ = o.roadmap_shortest_paths(nn, np.stack([eu, ev], 1), ew, sources[:n_cpu], threads=threads)
                cpu_s = time.perf_counter() - t0
                sp["cpu_port_sources_per_sec"] = n_cpu / cpu_s
                sp["cpu_port_threads"] = threads
                sp["distances_bit_equal_to_cpu_dijkstra"] = bool(np.array_equal(want, d_dist[:n_cpu].cpu().numpy()))
            out["config4_roadmap_shortest_paths_1024_sources"] = sp
            del d_dist, d_pred
        except Exception as e:
            out["config4_roadmap_shortest_paths_1024_sources"] = {"error": repr(e)}
        del d_ea, d_eb, d_eh, d_et, d_es, d_idx, d_nodes
    except Exception as e:  # an extra must never cost the headline line
        out["config4_prm_edges_100k_nodes_k10"] = {"error": repr(e)}
    return out


def parity_sample(mesh, a, b, got_hit, n_sample: int = 8192):
    """SURVEY §8d parity report on a sample of the timed batch: GPU booleans against the fp64 oracle, mismatches classified
    by the 1e-6 m clearance band (a motion is in-band if any state up to and including its first hit is)."""
    from oracle import oracle as o
    cores = os.cpu_count() or 1
    idx = np.linspace(0, len(a) - 1, min(n_sample, len(a))).astype(np.int64)
    w = o.Scene(mesh.vertices, mesh.triangles).check_motions(o.Robot(ARM, JOINTS), a[idx], b[idx], band=True, threads=cores)
    in_band = w["min_abs_clearance"] <= 1e-6
    mism = w["hit"] != got_hit[idx]
    return {"n_total": int(len(idx)), "n_mismatch_outside_band": int((mism & ~in_band).sum()), "n_in_band": int(in_band.sum()),
            "n_mismatch_in_band": int((mism & in_band).sum()), "band_m": 1e-6,
            "sample": f"{len(idx)} evenly spaced motions of the timed batch, booleans vs the fp64 oracle (oracle/mgodpl_oracle.hpp)"}


def cpu_baseline(mesh, a, b, use_ref: bool, budget_s: float = 20.0, cores: int | None = None):
    """The CPU implementation of the same step on the host cores (reported baseline, not the target)."""
    from oracle import oracle as o
    cores = cores or os.cpu_count() or 1
    m = len(a)
    if use_ref and o.have_ref():
        robot, scene = o.RefRobot(ARM, JOINTS), o.RefScene(mesh.vertices, mesh.triangles)

        def run(lo, hi, out):
            out.append(scene.check_motions(robot, a[lo:hi], b[lo:hi]))
        kind = "reference"
        what = ("reference's own collision_detection.cpp/RobotModel.cpp/... compiled unmodified (oracle/_ref); FCL is absent "
                "offline, fcl::collide answered by the oracle's exact SAT over an AABB tree")
    else:
        robot, scene = o.Robot(ARM, JOINTS), o.Scene(mesh.vertices, mesh.triangles)

        def run(lo, hi, out):
            r = scene.check_motions(robot, a[lo:hi], b[lo:hi])
            out.append((r["hit"], r["toi"]))
        kind = "port"
        what = "fp64 oracle port (FCL-equivalent CPU restatement; FCL not buildable offline)"
    # calibrate on a small slice, then size the sample for ~budget_s of wall time
    t0 = time.perf_counter()
    run(0, min(m, 500), [])
    per = (time.perf_counter() - t0) / min(m, 500)
    sample = int(min(m, max(2000, budget_s * cores / max(per, 1e-9) * 0.5)))
    # when the whole batch is less than the budget (about budget_s core-seconds of work), go over it several times
    reps = int(min(20, max(1, np.ceil(budget_s / max(per * sample, 1e-9)))))
    chunks = np.linspace(0, sample, cores + 1).astype(int)
    outs = [[] for _ in range(cores)]

    def run_reps(lo, hi, out):
        for _ in range(reps):
            run(lo, hi, out)
    threads = [threading.Thread(target=run_reps, args=(chunks[i], chunks[i + 1], outs[i])) for i in range(cores) if chunks[i + 1] > chunks[i]]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    hit = np.concatenate([x[0][0] for x in outs if x])
    toi = np.concatenate([x[0][1] for x in outs if x])
    # reference-equivalent state count of the sample, from the oracle's own step counts
    res = o.Scene(mesh.vertices, mesh.triangles).check_motions(o.Robot(ARM, JOINTS), a[:sample], b[:sample], threads=os.cpu_count() or 1)
    states = int(res["states_checked"].sum())
    assert np.array_equal(hit, res["hit"])
    return {"value": states * reps / dt, "unit": "states/s", "cores": cores, "kind": kind,
            "sample": f"first {sample} of {m} motions of the same batch x {reps} passes, {cores} host threads, {dt:.2f} s; {what}",
            "motions_per_sec": sample * reps / dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mgodpl_b200 as mg
    tree = mg.scenes.make_tree(**TREE)
    mesh = tree.collision_mesh(True)
    _, a, b = make_inputs(mg, 0)
    steps = max(1, args.steps)
    vals = []
    for k in range(args.warmup + steps):
        r = cpu_baseline(mesh, a, b, use_ref=True, budget_s=max(2.0, 60.0 / (args.warmup + steps)))
        if k >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    mps = float(np.mean([r["motions_per_sec"] for r in vals]))
    r = vals[-1]
    sample_n = int(r["sample"].split()[1])
    line = {"impl": "reference", "metric": "collision_checked_states_per_sec", "value": v, "unit": "states/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sample_n / mps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "motions_per_gpu": M_MOTIONS},
            "motions_per_sec": mps,
            "cpu_baseline": {"value": v, "unit": "states/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": v, "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="plain launches instead of replaying the device-timed step from CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
