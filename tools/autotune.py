#!/usr/bin/env python
"""Brute-force tile / launch-shape search on the GPU (the backend's version of the drivers' ``-a``).

For one kernel and one fusion grouping, time every candidate (tile shape x CUDA options) on the
large domain with the launcher's measurement loop and report Mpoints/s and effective GB/s.  The
best candidate per kernel can be written to ``absinthe_b200/tuned.json`` (what bench.py runs).

    python tools/autotune.py --kernel diffusion --assignment 0000111122223333 \
        --tiles 128x16x1,128x8x1,256x8x1 --budgets 110,220 --threads 256,512 --out gpurun_out/tune.json
"""

import argparse
import itertools
import json
import os
import statistics
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(kernel, assignment, shape, domain, cuda):
    from absinthe_b200 import programs as P
    from absinthe_b200.analysis import ceil_div
    seq = P.KERNELS[kernel][1]()
    groups = P.split_sequence(seq, assignment)
    # one tile shape (points) for every group, or a list of shapes, one per group
    shapes = [shape] * len(groups) if isinstance(shape[0], int) else list(shape)
    counts = [tuple(max(1, ceil_div(d, s)) for d, s in zip(domain, sh)) for sh in shapes]
    program = P.make_program(kernel, groups, counts, domain=domain,
                             variant="%s-%s" % (kernel, "_".join("x".join(map(str, sh)) for sh in shapes)))
    program["CUDA"] = dict(cuda)
    return program


def compile_one(args):
    from absinthe_b200.generator import build_variant
    try:
        return build_variant(make(*args)), None
    except Exception as exc:      # a candidate that does not fit is simply skipped
        return None, str(exc)[-300:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--assignment", default="", help="group index per stencil, e.g. 0000111122223333")
    ap.add_argument("--tiles", default="", help="comma list of tile shapes TXxTYxTZ")
    ap.add_argument("--from-tuned", action="store_true", help="time the variant recorded in tuned.json")
    ap.add_argument("--domain", default="1024x1024x80")
    ap.add_argument("--budgets", default="110", help="shared-memory budgets per CTA in KiB")
    ap.add_argument("--threads", default="256")
    ap.add_argument("--stages", default="2")
    ap.add_argument("--loads", default="tma")
    ap.add_argument("--subtiles", default="", help="optional comma list of BXxBYxBZ overrides")
    ap.add_argument("--extra", default="", help="JSON dict merged into every candidate's CUDA options")
    ap.add_argument("--sweep", default="", help="JSON dict option -> list of values (full product)")
    ap.add_argument("--runs", type=int, default=4)
    ap.add_argument("--out", default="")
    ap.add_argument("--save", action="store_true", help="write the winner into absinthe_b200/tuned.json")
    args = ap.parse_args()

    import torch
    from absinthe_b200 import programs as P
    from absinthe_b200.analysis import algorithmic_bytes_per_point, analyse
    from absinthe_b200.build import build_launcher
    from absinthe_b200.launcher import Variant, make_periodic
    build_launcher()

    domain = tuple(int(v) for v in args.domain.split("x"))
    if args.from_tuned:
        from absinthe_b200.analysis import ceil_div
        cfg = json.load(open(os.path.join(ROOT, "absinthe_b200", "tuned.json")))[args.kernel]
        args.assignment = "".join(str(a) for a in cfg["assignment"])
        args.tiles = "x".join(str(ceil_div(d, n)) for d, n in zip(domain, cfg["tiles"][0]))
        tuned_shapes = [tuple(ceil_div(d, n) for d, n in zip(domain, counts)) for counts in cfg["tiles"]]
        args.extra = json.dumps(cfg.get("cuda", {}))
        args.budgets = str(cfg.get("cuda", {}).get("SMEM_BUDGET", 110 * 1024) // 1024)
        args.threads = str(cfg.get("cuda", {}).get("THREADS", 256))
    assignment = [int(c) for c in args.assignment]
    shapes = [tuple(int(v) for v in t.split("x")) for t in args.tiles.split(",")]
    if args.from_tuned:
        shapes = [tuple(tuned_shapes)]          # the tuned variant: its own tile shape per group
    subtiles = [tuple(int(v) for v in t.split("x")) for t in args.subtiles.split(",") if t] or [None]
    extra = json.loads(args.extra) if args.extra else {}
    sweep = json.loads(args.sweep) if args.sweep else {}
    sweep_keys = sorted(sweep)
    candidates = []
    for shape, budget, threads, stages, load, sub in itertools.product(
            shapes, args.budgets.split(","), args.threads.split(","), args.stages.split(","),
            args.loads.split(","), subtiles):
        for combo in itertools.product(*[sweep[k] for k in sweep_keys]):
            cuda = {"SMEM_BUDGET": int(budget) * 1024, "THREADS": int(threads), "STAGES": int(stages),
                    "LOAD": load}
            if sub:
                cuda["SUBTILE"] = sub
            cuda.update(extra)
            cuda.update(dict(zip(sweep_keys, combo)))
            candidates.append((args.kernel, assignment, shape, domain, cuda))
    with ProcessPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        built = list(pool.map(compile_one, candidates))

    points = domain[0] * domain[1] * domain[2]
    results = []
    tensors = {}
    gen = torch.Generator(device="cuda").manual_seed(7)
    for cand, (paths, err) in zip(candidates, built):
        _, _, shape, _, cuda = cand
        if paths is None:
            print("skip", shape, cuda, err.splitlines()[-1] if err else "")
            continue
        v = Variant(paths[0], paths[1])
        dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
        for name in v.inputs + v.outputs:
            if name not in tensors:
                t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
                make_periodic(t, dims)
                tensors[name] = t
        v.bind([tensors[n] for n in v.inputs], [tensors[n] for n in v.outputs])
        total, halo = v.run(args.runs, flush=False)
        v.close()
        program = make(*cand)
        analyse(program)
        bpp = algorithmic_bytes_per_point(program)
        ms = statistics.median(total)
        hms = statistics.median(halo)
        row = {"tile": list(shape), "cuda": cuda, "total_ms": round(ms, 4), "halo_ms": round(hms, 4),
               "mpoints_s": round(points / ms / 1e3, 1), "GBps_total": round(bpp * points / ms / 1e6, 1),
               "GBps_kernels": round(bpp * points / (ms - hms) / 1e6, 1), "bytes_per_point": bpp}
        results.append(row)
        print(json.dumps(row), flush=True)
    results.sort(key=lambda r: r["total_ms"])
    if results:
        print("BEST", json.dumps(results[0]))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as handle:
            json.dump({"kernel": args.kernel, "assignment": assignment, "domain": list(domain),
                       "results": results}, handle, indent=1)
    if args.save and results:
        from absinthe_b200.analysis import ceil_div
        path = os.path.join(ROOT, "absinthe_b200", "tuned.json")
        tuned = json.load(open(path)) if os.path.exists(path) else {}
        best = results[0]
        per_group = best["tile"] if isinstance(best["tile"][0], (list, tuple)) else [best["tile"]] * (max(assignment) + 1)
        tuned[args.kernel] = {"assignment": assignment,
                              "tiles": [[max(1, ceil_div(d, s)) for d, s in zip(domain, sh)] for sh in per_group],
                              "cuda": best["cuda"], "measured": {k: best[k] for k in ("total_ms", "GBps_total")}}
        with open(path, "w") as handle:
            json.dump(tuned, handle, indent=1)


if __name__ == "__main__":
    main()
