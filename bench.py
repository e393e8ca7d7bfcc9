#!/usr/bin/env python
"""Benchmark of the hot path: fused stencil sequences on a large synthetic domain.

``python bench.py --gpus N --steps K --warmup W`` -- one rank per GPU (torchrun for N > 1).
A *step* is one pass of the diffusion sequence and one pass of the fast-waves sequence
(BASELINE.json configs[4]) over a 1024 x 1024 x 80 domain per GPU, each executed as its tuned
fused/tiled variant through the C-ABI launcher, periodic halo updates included.  Printed metric:
Mpoints/s, a point being one grid point taken through one whole sequence.

``--impl reference`` times the reference's own OpenMP build (oracle/_ref, built in the
development container) of the same variants on all host cores, on the full 1024 x 1024 x 80 domain
(the ``cpu_baseline`` leg of the GPU arm uses a 1024 x 1024 x 8 slab of it to stay within seconds).
``--check`` runs one step of the bench variants (through the slab exchange when N > 1) and compares
windows at the domain corners and at the j-seams of the slabs with the CPU oracle.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DOMAIN = (1024, 1024, 80)
CPU_SLAB = (1024, 1024, 8)          # bounded sample for the CPU arms (same per-point work)
TUNED = os.path.join(ROOT, "absinthe_b200", "tuned.json")
TRAFFIC = os.path.join(ROOT, "profiles", "traffic.json")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def workload_programs(domain=DOMAIN):
    """The two variants of the step: (name, program) with tuned groupings / tile counts."""
    from absinthe_b200 import programs as P
    tuned = {}
    if os.path.exists(TUNED):
        tuned = json.load(open(TUNED))
    out = []
    for kernel in ("diffusion", "fastwaves"):
        seq = P.KERNELS[kernel][1]()
        cfg = tuned.get(kernel, DEFAULT_VARIANTS[kernel])
        groups = P.split_sequence(seq, cfg["assignment"])
        program = P.make_program(kernel, groups, [tuple(t) for t in cfg["tiles"]], domain=domain,
                                 variant="%s-%s" % (kernel, "-".join(str(a) for a in cfg["assignment"])))
        if cfg.get("cuda"):
            program["CUDA"] = dict(cfg["cuda"])
        out.append((kernel, program))
    return out


REF_RUNS = 6                        # printed runs compiled into the reference binaries (2 x RUNS executions)


def extra_programs():
    """Variants timed for the ``extras`` block (outside ``value``): advection at the bench domain, auto-tuned
    and with the HAND-style one-tile-per-plane tiling (ref advection.py:145-160), and the three sequences on the
    reference's default 64 x 64 x 60 domain (BASELINE.json configs[0], [2])."""
    from absinthe_b200 import programs as P
    tuned = json.load(open(TUNED)) if os.path.exists(TUNED) else {}
    out = []
    adv = P.KERNELS["advection"][1]()
    cfg = tuned.get("advection", {"assignment": [0] * 8, "tiles": [[2, 128, 80]], "cuda": {}})
    auto = P.make_program("advection", P.split_sequence(adv, cfg["assignment"]), [tuple(t) for t in cfg["tiles"]],
                          domain=DOMAIN, variant="advection-auto")
    auto["CUDA"] = dict(cfg.get("cuda", {}))
    hand = P.make_program("advection", [adv], [(1, 1, DOMAIN[2])], domain=DOMAIN, variant="advection-hand")
    hand["CUDA"] = dict(cfg.get("cuda", {}))
    out += [("advection:auto", auto), ("advection:hand(1,1,Z)", hand)]
    small = {"INLINE": "all", "FUSE_HALO": True, "BLOCK": (2, 1), "THREADS": 256}
    for kernel, groups, cuda in (
            ("diffusion", None, dict(small, SHUFFLE=True)),
            ("advection", None, small),
            ("fastwaves", "split", dict(small, BLOCK=(1, 2)))):
        seq = P.KERNELS[kernel][1]()
        groups = [seq[:6], seq[6:]] if groups == "split" else [seq]
        program = P.make_program(kernel, groups, [(2, 16, 5)] * len(groups), variant=kernel + "-64x64x60")
        program["CUDA"] = dict(cuda)
        out.append(("small:" + kernel, program))
    # BASELINE.json configs[1]: a few programs of the training sweeps, every tile executed (20 per SM)
    for family, variant in (("fitcache", "PT12"), ("fitcache", "PT20"), ("fitddr", "IN1BD0"), ("fitddr", "IN3BD0")):
        out.append(("training:%s-%s" % (family, variant), P.training_program(family, variant, (50, 5, 8), stride=1)))
    return out


# untuned fall-backs: diffusion one group per field, fast waves the HAND/AUTO two-group split
DEFAULT_VARIANTS = {
    "diffusion": {"assignment": [0] * 4 + [1] * 4 + [2] * 4 + [3] * 4, "tiles": [[8, 64, 80]] * 4},
    "fastwaves": {"assignment": [0] * 6 + [1] * 3, "tiles": [[8, 64, 10]] * 2},
}


def scale_tiles(program, domain):
    """The same tile *shapes* on another domain (tile counts follow the extents)."""
    from absinthe_b200 import programs as P
    from absinthe_b200.analysis import ceil_div
    q = P.clone(program)
    old = (q["X"], q["Y"], q["Z"])
    q["X"], q["Y"], q["Z"] = domain
    for outer in q["TILING"]["GROUPS"]:
        for inner in outer["GROUPS"]:
            for dim, o, n in zip("XYZ", old, domain):
                shape = ceil_div(o, inner["N" + dim])
                inner["N" + dim] = max(1, ceil_div(n, shape))
    return q


def template_of(program):
    return "template_training.cu" if program.get("TRAINING") else "template_tiling.cu"


def prebuild_variants():
    from absinthe_b200.generator import build_variant
    from absinthe_b200 import programs as P
    for _, program in workload_programs() + extra_programs():
        build_variant(P.clone(program), template_of(program))


def prebuild():
    """Compile everything the default run needs (called from __graft_entry__.build)."""
    prebuild_variants()
    try:
        from oracle import build_ref
        if build_ref.available():
            for name, program in workload_programs():
                for prefix, domain in (("bench_", CPU_SLAB), ("benchfull_", DOMAIN)):
                    sample = scale_tiles(program, domain)
                    sample["RUNS"] = REF_RUNS
                    sample["VARIANT"] = program["VARIANT"]
                    build_ref.build_binary(sample, prefix + name)
    except Exception as exc:
        print("reference build skipped:", exc)


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc, self.window = index, [], None, [None, None, None]

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def mark(self, which):
        """Host time stamps: 0 = start of the timed region, 1 = after its final synchronize, 2 = warm-up start."""
        self.window[which] = time.monotonic()

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        return self.summary(*self.window)

    def summary(self, lo, hi, warm=None):
        """Clock statistics of the samples between two ``time.monotonic()`` stamps."""
        sm, reasons, sm_max = [], set(), None
        timed = [row for t, row in self.rows if lo is not None and hi is not None and lo <= t <= hi + 0.02]
        # a timed region shorter than a few sampling periods holds too few samples: then the warm-up steps
        # that ran right before it (same kernels, same load) are counted too, and the line says so
        window = "timed region" if len(timed) >= 3 else "warm-up + timed region"
        for row in (timed if len(timed) >= 3 else [row for t, row in self.rows if (warm is None or t >= warm)
                                                    and (hi is None or t <= hi + 0.02)]):
            try:
                sm.append(float(row[0]))
                sm_max = float(row[1])
            except (ValueError, IndexError):
                continue
            for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                  row[3:7]):
                if cell == "Active":
                    reasons.add(name)
        busy = [c for c in sm if sm_max and c > 0.3 * sm_max] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": sm_max,
                "reasons": sorted(reasons), "samples": len(sm), "window": window, "period_ms": 20}


def cpu_reference_time(name, steps, warmup, prefix="bench_"):
    """Run the prebuilt reference binary of ``name``; returns (median total ms, samples, threads)."""
    exe = os.path.join(REF_DIR, prefix + name)
    if not os.path.exists(exe):
        return None
    threads = os.cpu_count() or 1
    env = dict(os.environ, OMP_PLACES="cores", OMP_PROC_BIND="close", OMP_NUM_THREADS=str(threads),
               OMP_STACKSIZE="512M")
    proc = subprocess.run("ulimit -s unlimited; exec " + exe, shell=True, env=env, capture_output=True, text=True,
                          timeout=1500)
    totals = [float(line.split(":")[1]) for line in proc.stdout.splitlines()
              if line.startswith("   - total time")]
    if not totals:
        return None
    used = totals[min(warmup, len(totals) - 1):][:max(1, steps)]
    return statistics.median(used), len(used), threads


def cpu_port_time(name, program):
    """Fallback CPU baseline: the sequential oracle restatement on one core."""
    from oracle.reference_oracle import Oracle
    import numpy as np
    oracle = Oracle(program, fast_math=True)
    rng = np.random.default_rng(0)
    inputs = [oracle.make_periodic(rng.uniform(0.5, 1.5, oracle.shape)) for _ in oracle.inputs]
    t0 = time.perf_counter()
    oracle.run(inputs)
    return (time.perf_counter() - t0) * 1e3


def cpu_baseline(steps=3, warmup=1, full=False):
    """Mpoints/s of the CPU arm; prefers the real reference build.  ``full``: the whole bench domain (the
    ``--impl reference`` arm), else a bounded 1024 x 1024 x 8 slab of it (same per-point work)."""
    from absinthe_b200 import programs as P
    domain = DOMAIN if full else CPU_SLAB
    points = domain[0] * domain[1] * domain[2]
    total_ms, kind, threads, detail, timed = 0.0, "reference", 1, {}, steps
    for name, program in workload_programs():
        got = cpu_reference_time(name, steps, warmup, "benchfull_" if full else "bench_")
        if got is None and full:                 # no full-size binary shipped: the slab, and the line says so
            return cpu_baseline(steps, warmup, full=False)
        if got is None:
            kind = "port"
            break
        total_ms += got[0]
        timed, threads = got[1], got[2]      # the binary holds RUNS = 6: at most 6 - warmup timed runs
        detail[name] = got[0]
    if kind == "port":
        total_ms, threads, detail = 0.0, 1, {}
        domain = (256, 256, 8)
        points = domain[0] * domain[1] * domain[2]
        for name, program in workload_programs():
            ms = cpu_port_time(name, scale_tiles(program, domain))
            total_ms += ms
            detail[name] = ms
        sample = "sequential oracle port, %dx%dx%d slab, one pass per sequence" % domain
    else:
        sample = ("reference OpenMP build (g++ -O3 -ffast-math -fopenmp) of the same variants on %s, "
                  "median of %d timed runs per sequence" % (
                      "the full %dx%dx%d domain" % domain if domain == DOMAIN else "a %dx%dx%d slab" % domain, timed))
    value = 2 * points / (total_ms * 1e-3) / 1e6
    return {"value": value, "unit": "Mpoints/s", "cores": threads, "kind": kind, "sample": sample,
            "sample_domain": list(domain), "ms_per_sequence": detail}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(steps=args.steps, warmup=args.warmup, full=True)
    points = base["sample_domain"][0] * base["sample_domain"][1] * base["sample_domain"][2]
    config = workload_config(args.gpus)
    config["reference_sample"] = base["sample"]
    config["same_domain_as_gpu_arm"] = base["sample_domain"] == list(DOMAIN)
    line = {
        "impl": "reference", "metric": "fused_sequence_throughput", "value": base["value"], "unit": "Mpoints/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 2 * points / base["value"] / 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config, "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def _available_host_bytes():
    """MemAvailable of the box (0 when /proc/meminfo cannot be read: the caller then stays frugal)."""
    try:
        with open("/proc/meminfo") as handle:
            for line in handle:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def workload_config(gpus, exchange=None):
    return {"workload": "diffusion+fastwaves fused sequences, %dx%dx%d fp64 per GPU, halo 3, periodic" % DOMAIN,
            "domain_per_gpu": list(DOMAIN), "decomposition": "j-slabs x%d" % gpus, "exchange": exchange,
            "l2": "inputs (>6 GB per sequence) exceed the 126 MB L2; no flush needed between steps"}


# ---------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from absinthe_b200 import programs as P
    from absinthe_b200.analysis import algorithmic_bytes_per_point
    from absinthe_b200.build import build_launcher
    from absinthe_b200.launcher import Variant, make_periodic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    sampler = ClockSampler(local)                  # started early: nvidia-smi takes a while to come up
    sampler.start()
    if local == 0:                                 # one rank builds, the others wait for the cache
        build_launcher()
        prebuild_variants()
    if world > 1:
        dist.barrier()

    from absinthe_b200.distributed import P2PExchange, SlabExchange
    Exchange = P2PExchange if args.exchange == "p2p" else SlabExchange
    variants = []
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    for name, program in workload_programs():
        v = Variant.from_program(P.clone(program), device=local)
        dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
        ins = []
        for _ in v.inputs:
            t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
            make_periodic(t, dims)
            ins.append(t)
        outs = [v.empty_field() for _ in v.outputs]
        v.bind(ins, outs)
        exchange = None
        if world > 1:
            try:
                exchange = Exchange(v, world, rank)
            except RuntimeError as exc:                # raised on every rank alike (see P2PExchange)
                if rank == 0:
                    print("falling back to the NCCL exchange:", exc, file=sys.stderr)
                exchange, args.exchange = SlabExchange(v, world, rank), "nccl (p2p unavailable)"
        analysed = P.clone(program)
        from absinthe_b200.analysis import analyse
        analyse(analysed)
        groups = [g for o in analysed["TILING"]["GROUPS"] for g in o["GROUPS"]]
        variants.append({"name": name, "v": v, "ins": ins, "outs": outs, "exchange": exchange,
                         "bytes_per_point": algorithmic_bytes_per_point(analysed),
                         "group_bytes": [8 * (len(g["INPUTS"]) + len(g["OUTPUTS"])) for g in groups],
                         "group_names": ["%s:g%d(%s..%s)" % (name, n + 1, g["STENCILS"][0]["NAME"],
                                                             g["STENCILS"][-1]["NAME"]) for n, g in enumerate(groups)]})
    points = DOMAIN[0] * DOMAIN[1] * DOMAIN[2]

    def step(record=None):
        for item in variants:
            v, exchange = item["v"], item["exchange"]
            for g in range(1, v.n_groups + 1):
                if exchange is not None:
                    exchange.wait(g)              # last step's exchange still reads these outputs
                    if g > 1:
                        exchange.wait(g - 1)      # WAIT: this group reads the halos of the previous one
                if record is not None:
                    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    a.record()
                if exchange is not None:
                    exchange.apply(g)             # COMP + PUT (overlapped with the next kernels, or fused into this one)
                else:
                    v.apply_group(g, with_halo=False)
                if record is not None:
                    b.record()
                if exchange is None:
                    v.halo_group(g)
                if record is not None:
                    c.record()
                    record.append((item["name"], g, a, b, c))

    # the sampler was started before the variants were built; make sure nvidia-smi is delivering samples
    # before the GPU work begins (at most 3 s), then W (>= 3) warm-up steps
    wait_t0 = time.monotonic()
    while not sampler.rows and time.monotonic() - wait_t0 < 3.0:
        time.sleep(0.01)
    warm = max(args.warmup, 3)
    sampler.mark(2)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    records = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler.mark(0)
    start.record()
    for _ in range(args.steps):
        step(records)
    for item in variants:                         # drain the exchanges still in flight
        if item["exchange"] is not None:
            for g in range(1, item["v"].n_groups + 1):
                item["exchange"].wait(g)
    stop.record()
    torch.cuda.synchronize()
    sampler.mark(1)
    elapsed_ms = start.elapsed_time(stop)
    clocks = sampler.summary(*sampler.window)

    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = 2 * points * world / (ms_per_step * 1e-3) / 1e6

    # sustained: the same step back to back for >= `--sustain` seconds (the board settles at its power cap); the
    # last batch is reported, with per-kernel events.  The batch count is fixed up front: every rank runs the same.
    sustained, sustained_records = None, []
    if args.sustain > 0:
        per_batch = 25
        n_batches = max(2, int(args.sustain * 1e3 / (ms_per_step * per_batch)) + 1)
        t_begin = time.monotonic()
        for batch in range(n_batches):
            last = batch == n_batches - 1
            if last:
                torch.cuda.synchronize()
                t_last = time.monotonic()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(per_batch):
                step(sustained_records if last else None)
            if last:
                for item in variants:
                    if item["exchange"] is not None:
                        for g in range(1, item["v"].n_groups + 1):
                            item["exchange"].wait(g)
            b.record()
        torch.cuda.synchronize()
        t_end = time.monotonic()
        sustained_ms = a.elapsed_time(b) / per_batch
        if world > 1:
            t = torch.tensor([sustained_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sustained_ms = float(t.item())
        sustained = {"ms_per_step": round(sustained_ms, 4), "seconds": round(t_end - t_begin, 2),
                     "steps": n_batches * per_batch,
                     "value": round(2 * points * world / (sustained_ms * 1e-3) / 1e6, 1), "unit": "Mpoints/s",
                     "clocks": sampler.summary(t_last, t_end)}
    sampler.stop()

    # per-kernel durations from the events of the timed region
    kernel_ms, halo_ms = {}, {}
    for name, g, a, b, c in records:
        kernel_ms.setdefault((name, g), []).append(a.elapsed_time(b))
        halo_ms.setdefault((name, g), []).append(b.elapsed_time(c))
    # time between the end of the previous group kernel and the start of this one (host enqueue gaps, and on
    # several GPUs the part of the halo exchange the kernels could not hide)
    gap_ms = {}
    for prev, this in zip(records, records[1:]):
        gap_ms.setdefault((this[0], this[1]), []).append(prev[3].elapsed_time(this[2]))
    breakdown = {}
    dominant = None
    for item in variants:
        for g in range(1, item["v"].n_groups + 1):
            k = statistics.mean(kernel_ms[(item["name"], g)])
            h = statistics.mean(halo_ms[(item["name"], g)])
            gbs = item["group_bytes"][g - 1] * points / (k * 1e-3) / 1e9
            label = item["group_names"][g - 1]
            breakdown[label] = {"kernel_ms": round(k, 4), "halo_ms": round(h, 4), "GBps": round(gbs, 1),
                                "bytes_per_point": item["group_bytes"][g - 1],
                                "gap_before_ms": round(statistics.mean(gap_ms.get((item["name"], g), [0.0])), 4)}
            if dominant is None or k > dominant[1]:
                dominant = (label, k, gbs)
    peaks = {}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak, peak_kind = (peaks["hbm_gbs"], "measured copy bandwidth (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback of B200_PROFILING.md")
    traffic = None
    if os.path.exists(TRAFFIC):
        traffic = json.load(open(TRAFFIC)).get(dominant[0])
    if sustained is not None:
        sustained["step_frac"] = round(sum(item["bytes_per_point"] for item in variants) * points
                                       / (sustained["ms_per_step"] * 1e-3) / 1e9 / peak, 4)
        hot = {}
        for name, g, a, b, c in sustained_records:
            hot.setdefault((name, g), []).append(a.elapsed_time(b))
        kernels = {}
        for item in variants:
            for g in range(1, item["v"].n_groups + 1):
                k = statistics.mean(hot[(item["name"], g)])
                gbs = item["group_bytes"][g - 1] * points / (k * 1e-3) / 1e9
                kernels[item["group_names"][g - 1]] = {"kernel_ms": round(k, 4), "GBps": round(gbs, 1),
                                                       "frac": round(gbs / peak, 4)}
        sustained["kernels"] = kernels
        sustained["frac"] = kernels[dominant[0]]["frac"]          # the roofline kernel, in the heated state
    step_bytes = sum(item["bytes_per_point"] for item in variants) * points
    roofline = {"bound": "hbm", "kernel": dominant[0], "achieved": round(dominant[2], 1), "peak": peak,
                "unit": "GB/s", "frac": round(dominant[2] / peak, 4), "traffic": traffic, "peak_kind": peak_kind,
                "step_GBps": round(step_bytes * world / (ms_per_step * 1e-3) / 1e9, 1),
                "step_frac": round(step_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4),
                "frac_of_8TBps": round(dominant[2] / 8000.0, 4)}

    extras = None
    if world == 1 and not args.no_extras:
        try:
            extras = measure_extras(peak)
        except Exception as exc:                       # extras never cost the headline line
            extras = {"error": str(exc)[:300]}

    # end to end through the public API with HOST buffers (pinned), copies inside the timed region
    launches = sum(item["v"].launches for item in variants) * args.steps
    e2e_steps = max(1, min(args.steps, 10))
    host = []
    constants = {}
    nfields = sum(len(item["v"].inputs) + 2 * len(item["v"].outputs) for item in variants)
    ranks_on_box = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    # two steps of a sequence in flight need their own host output buffers; fall back to one step in
    # flight when the box could not pin that much memory for all its ranks
    depth = 2 if _available_host_bytes() > 2 * ranks_on_box * nfields * variants[0]["v"].field_bytes else 1
    for (name, program), item in zip(workload_programs(), variants):
        v = item["v"]
        # the program's CONSTANTS (mask; wgtfac, hhl, rho, dzdx, dzdy) stay resident on the device as
        # bound; every step uploads the fields that change from step to step and downloads all outputs
        const = [n for n in v.inputs if n in program.get("CONSTANTS", [])]
        constants[name] = const
        hin = [None if n in const else t.cpu().pin_memory() for n, t in zip(v.inputs, item["ins"])]
        hout = [[torch.empty(v.shape, dtype=torch.float64).pin_memory() for _ in v.outputs]
                for _ in range(depth)]
        host.append((v, hin, hout, torch.cuda.Stream()))   # own compute stream: no head-of-line blocking
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    # software pipeline: every sequence keeps `depth` steps in flight.  The upload of step n+1 starts when
    # the kernels of step n have consumed the device inputs and overlaps the download of step n
    # (full-duplex PCIe); every step's outputs are waited for on the host (wait_host) before its host
    # buffers are handed to a later step.
    for n in range(min(depth, e2e_steps)):
        for v, hin, hout, stream in host:
            v.apply_host(hin, hout[n % depth], stream=stream, wait=False)
    for n in range(e2e_steps):
        for v, hin, hout, stream in host:
            v.wait_host()                        # step n's outputs of the sequence are on the host
            if n + depth < e2e_steps:
                v.apply_host(hin, hout[n % depth], stream=stream, wait=False)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    # bytes that cross the bus per uploaded field: its interior, or the whole padded field
    up_bytes = points * 8 if args.e2e_upload == "interior" else host[0][0].field_bytes
    h2d = sum(sum(1 for t in hin if t is not None) for _, hin, _, _ in host) * up_bytes
    d2h = sum(len(hout[0]) for _, _, hout, _ in host) * host[0][0].field_bytes
    e2e = {"value": round(2 * points * world / (e2e_ms * 1e-3) / 1e6, 1), "unit": "Mpoints/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms, 2),
           "steps": e2e_steps, "constants_resident": constants, "steps_in_flight": depth, "upload": args.e2e_upload,
           "api": "Variant.apply_host(wait=False) + wait_host per sequence and step "
                  "(absinthe_variant_apply_host_async / absinthe_variant_wait_host)"}

    if rank == 0:
        line = {
            "metric": "fused_sequence_throughput", "value": round(value, 1), "unit": "Mpoints/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world, args.exchange if world > 1 else None), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "breakdown": breakdown, "sustained": sustained,
            "extras": extras,
            "variants": {item["name"]: workload_programs()[n][1]["VARIANT"] for n, item in enumerate(variants)},
        }
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"] = cpu_baseline()
            except Exception as exc:
                line["cpu_baseline"] = {"error": str(exc)[:200]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_extras(peak):
    """Driver-visible numbers for BASELINE.json configs[0..3], outside ``value`` (see ``extra_programs``)."""
    import torch
    from absinthe_b200 import programs as P
    from absinthe_b200.analysis import algorithmic_bytes_per_point, analyse
    from absinthe_b200.launcher import Variant, make_periodic
    out = {}
    gen = torch.Generator(device="cuda").manual_seed(99)
    for label, program in extra_programs():
        v = Variant.from_program(P.clone(program), template_of(program))
        dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
        ins = []
        for _ in v.inputs:
            t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
            make_periodic(t, dims)
            ins.append(t)
        outs = [v.empty_field() for _ in v.outputs]
        v.bind(ins, outs)
        if label.startswith("training:"):       # the training protocol of the launcher (ref template_training.cpp:276-311)
            total, _ = v.run(12, flush=False, training=True)
            ms = statistics.median(total)
            tiles = 2 * 2 * 5 * program["TRAINING"]["CORES"]
            out[label] = {"domain": [v.X, v.Y, v.Z], "tile": [v.X // 2, v.Y // 2, v.Z // (5 * program["TRAINING"]["CORES"])],
                          "tiles": tiles, "ms": round(ms, 4), "ns_per_tile_and_sm": round(ms * 1e6 / (tiles / 148.0), 1),
                          "mpoints_s": round(v.points / ms / 1e3, 1)}
            v.close()
            del ins, outs
            torch.cuda.empty_cache()
            continue
        small = label.startswith("small:")
        reps = 300 if small else 12
        for _ in range(3):
            v.apply(graph=small)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            v.apply(graph=small)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        row = {"domain": [v.X, v.Y, v.Z], "variant": program["VARIANT"], "launches": v.launches,
               "mpoints_s": round(v.points / ms / 1e3, 1)}
        if small:                                   # the whole domain sits in L2: this is a launch latency
            row["graph_us"] = round(ms * 1e3, 2)
        else:
            bpp = algorithmic_bytes_per_point(analyse(P.clone(program)))
            gbs = bpp * v.points / (ms * 1e-3) / 1e9
            row.update(ms=round(ms, 4), bytes_per_point=bpp, GBps=round(gbs, 1), frac=round(gbs / peak, 4))
        out[label] = row
        v.close()
        del ins, outs
        torch.cuda.empty_cache()
    return out


def run_check_mode(args):
    """``--check``: one step of the bench variants on every rank, windows (corners, middle, both j-seams of the
    slab) against the CPU oracle.  The oracle is the checker here; nothing of this is timed."""
    import torch
    import torch.distributed as dist
    from absinthe_b200.build import build_launcher
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from seam_check import run_check
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if local == 0:
        build_launcher()
        prebuild_variants()
    if world > 1:
        dist.barrier()
    report = run_check(workload_programs(), world, rank, local, exchange_kind=args.exchange)
    reports = [report]
    if world > 1:
        reports = [None] * world
        dist.all_gather_object(reports, report)
    if rank == 0:
        emit({"check": "ok" if all(r["ok"] for r in reports) else "FAILED", "n_gpus": world,
              "config": workload_config(world), "windows_per_rank": len(report["windows"]),
              "failed": [w for r in reports for w in r["windows"] if not w["ok"]],
              "rows_compared": sum(w["rows"] for r in reports for w in r["windows"])})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not report["ok"]:
        sys.exit(1)


RESULT = sys.stdout


def claim_stdout():
    """Keep stdout for the one JSON line: whatever libraries write to file descriptor 1 (NCCL prints a
    version banner there under torchrun, whatever NCCL_DEBUG_FILE says) is sent to stderr instead."""
    global RESULT
    sys.stdout.flush()
    RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    RESULT.write(json.dumps(line) + "\n")
    RESULT.flush()


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=100)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="absinthe_b200", choices=["absinthe_b200", "reference"])
    parser.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    parser.add_argument("--no-extras", action="store_true", help="skip the extras block (advection, 64x64x60 latencies)")
    parser.add_argument("--sustain", type=float, default=2.0,
                        help="seconds of back-to-back steps after the timed region for the sustained sub-record (0 = off)")
    parser.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                        help="halo exchange of the multi-GPU run: peer-to-peer windows inside the launcher "
                             "(absinthe_variant_put_group / _wait_group) or pack + NCCL send/recv + unpack")
    parser.add_argument("--check", action="store_true",
                        help="run one step of the bench variants and compare corner and seam windows with the CPU oracle")
    parser.add_argument("--e2e-upload", default="full", choices=["full", "interior"],
                        help="end-to-end leg: upload whole padded fields, or only the interiors of the host inputs "
                             "(the library rebuilds their periodic halos on the device; measured: same step time)")
    args = parser.parse_args()
    claim_stdout()
    if args.check:
        run_check_mode(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
