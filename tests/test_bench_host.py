"""Host-side logic of bench.py that needs no GPU: clock-sample windows, the stdout contract, workloads."""

import json
import os
import subprocess
import sys
import time

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IDLE = ["300", "1965", "100", "Not Active", "Not Active", "Not Active", "Not Active"]
BUSY = ["1900", "1965", "950", "Not Active", "Not Active", "Not Active", "Active"]
FULL = ["1965", "1965", "900", "Not Active", "Not Active", "Not Active", "Not Active"]


def sampler(rows, window):
    s = bench.ClockSampler(0)
    s.join = lambda timeout=None: None          # never started: nothing to join
    s.rows, s.window = rows, window
    return s.stop()


def test_clock_samples_of_the_timed_region_are_preferred():
    now = time.monotonic()
    rows = [(now - 1.0, IDLE), (now + 0.01, BUSY), (now + 0.11, FULL), (now + 0.13, FULL), (now + 0.15, BUSY)]
    got = sampler(rows, [now + 0.1, now + 0.2, now])
    assert got["window"] == "timed region" and got["samples"] == 3
    assert got["sm_mhz"] == 1965.0 and got["sm_max_mhz"] == 1965.0 and got["reasons"] == ["sw_power_cap"]


def test_short_timed_region_falls_back_to_the_warm_up_samples():
    now = time.monotonic()
    rows = [(now - 1.0, IDLE), (now + 0.01, BUSY), (now + 0.03, FULL), (now + 0.11, FULL)]
    got = sampler(rows, [now + 0.1, now + 0.12, now])
    assert got["window"] == "warm-up + timed region" and got["samples"] == 3      # the idle sample before warm-up is out
    assert got["sm_mhz"] == 1965.0
    assert sampler([], [now, now + 1, now - 1])["sm_mhz"] is None


def test_stdout_carries_only_the_json_line():
    """Whatever a library writes to file descriptor 1 (NCCL's banner) must not end up next to the JSON line."""
    code = ("import bench, os; bench.claim_stdout(); os.system('echo banner'); print('python noise'); "
            "bench.emit({'metric': 'x'})")
    proc = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0
    assert json.loads(proc.stdout) == {"metric": "x"}
    assert "banner" in proc.stderr and "python noise" in proc.stderr


def test_workload_is_the_named_configuration():
    programs = dict(bench.workload_programs())
    assert sorted(programs) == ["diffusion", "fastwaves"]
    for program in programs.values():
        assert (program["X"], program["Y"], program["Z"]) == (1024, 1024, 80)
        assert (program["HX"], program["HY"], program["HZ"]) == (3, 3, 3)
    config = bench.workload_config(8)
    assert config["domain_per_gpu"] == [1024, 1024, 80] and config["decomposition"] == "j-slabs x8"
    # the CPU arm runs the same variants (tile shapes) on a bounded slab
    slab = bench.scale_tiles(programs["fastwaves"], bench.CPU_SLAB)
    assert (slab["X"], slab["Y"], slab["Z"]) == bench.CPU_SLAB
    for outer_full, outer_slab in zip(programs["fastwaves"]["TILING"]["GROUPS"], slab["TILING"]["GROUPS"]):
        a, b = outer_full["GROUPS"][0], outer_slab["GROUPS"][0]
        assert (a["NX"], a["NY"]) == (b["NX"], b["NY"]) and b["NZ"] == 1


def test_extras_cover_the_other_baseline_configs():
    """BASELINE.json configs[0..3] get driver-visible numbers in the `extras` block of the bench line."""
    labels = [label for label, _ in bench.extra_programs()]
    assert "advection:auto" in labels and "advection:hand(1,1,Z)" in labels
    assert {"small:diffusion", "small:advection", "small:fastwaves"} <= set(labels)
    assert sum(label.startswith("training:") for label in labels) == 4
    by_label = dict(bench.extra_programs())
    hand = by_label["advection:hand(1,1,Z)"]["TILING"]["GROUPS"][0]["GROUPS"][0]
    assert (hand["NX"], hand["NY"], hand["NZ"]) == (1, 1, 80)            # ref advection.py:145-160, one tile per plane
    assert by_label["small:fastwaves"]["X"] == 64 and by_label["training:fitddr-IN3BD0"]["TRAINING"]["STRIDE"] == 1
    assert bench.template_of(by_label["training:fitcache-PT12"]) == "template_training.cu"


def test_reference_arm_names_its_sample(monkeypatch):
    """--impl reference states which domain the reference binary ran on (the full one when it is shipped)."""
    monkeypatch.setattr(bench, "cpu_reference_time", lambda name, steps, warmup, prefix="bench_": (100.0, 3, 16))
    full = bench.cpu_baseline(steps=3, warmup=1, full=True)
    assert full["sample_domain"] == [1024, 1024, 80] and "full 1024x1024x80 domain" in full["sample"]
    slab = bench.cpu_baseline(steps=3, warmup=1)
    assert slab["sample_domain"] == [1024, 1024, 8] and slab["kind"] == "reference" and slab["cores"] == 16
    # value: two sequences over the sample's points per summed milliseconds
    assert abs(full["value"] - 2 * 1024 * 1024 * 80 / 0.2 / 1e6) < 1e-6
