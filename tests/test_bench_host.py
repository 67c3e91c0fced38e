"""Host-side pieces of bench.py that run without a GPU: the reference-equivalent state count, the clock sampler's
pause / sample-between-calls protocol, and the reference arm's JSON line (the contract the driver parses)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_equivalent_states_follow_the_sequential_loop():
    """collision_detection.cpp:94-103 checks t = i / n_steps for i = 0 .. n_steps and returns at the first hit (toi = that t):
    a motion that first collides at step k costs k + 1 checks, a free one n_steps + 1, a degenerate one (n_steps = 0) one."""
    hit = np.array([True, False, True, True, False, True])
    toi = np.array([0.5, np.nan, 0.0, 1.0, np.nan, 0.0])
    n_steps = np.array([4, 7, 3, 5, 0, 0])
    assert bench.ref_equivalent_states(hit, toi, n_steps).tolist() == [3, 8, 1, 6, 1, 1]
    # every representable first-hit step of a long motion survives the round trip through toi = k / n
    n = 997
    k = np.arange(n + 1)
    assert np.array_equal(bench.ref_equivalent_states(np.ones(n + 1, bool), k / n, np.full(n + 1, n)), k + 1)


def test_clock_sampler_pause_and_between_call_samples():
    s = bench.ClockSampler(0)          # thread not started: no NVML here
    s.sm, s.mask = [1500, 1600], 0x8   # what a poller had gathered before the pause
    s._query = lambda: (1965, 0x4)
    s.max_sm = 1965
    s.pause()
    assert s.paused.is_set() and s.sm == [] and s.mask == 0
    for _ in range(3):
        s.sample_now()
    out = s.summary()
    assert out["sm_mhz"] == 1965 and out["samples"] == 3 and out["reasons"] == ["sw_power_cap"] and out["sm_max_mhz"] == 1965
    s._query = None                    # NVML unavailable: sampling from the caller is a no-op, the summary says so
    s.sm, s.mask = [], 0
    s.sample_now()
    assert s.summary()["reasons"] == ["unsampled"]

    def broken():
        raise RuntimeError("driver went away")
    s._query = broken
    s.sample_now()                     # never propagates into the timed loop


def test_byte_accounting_is_the_surveys():
    assert bench.B_MOTION == 152.125 and bench.B_SCENE_PER_TRIANGLE == 112


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "collision_checked_states_per_sec" and line["unit"] == "states/s"
    assert line["higher_is_better"] is True and line["value"] > 1e5 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["config"]["workload"] == bench.WORKLOAD          # identical string in both arms (same_config)
    assert line["e2e"] == {"value": line["value"], "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "motions" in cb["sample"]
    assert line["vs_baseline"] is None and line["dtype"] == "f64"
