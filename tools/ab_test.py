#!/usr/bin/env python
"""Interleaved A/B timing of CUDA option sets on one tuned variant (removes warm-up / power-state order bias).

    ab_test.py [--prebuild] <kernel> '<json options A>' '<json options B>' ...

An option set may hold "_shapes": [[x, y, z], ...] -- tile shapes in points, one per group (or one for all) --
which replaces the tuned tile counts, "_assignment": [group index per stencil] which replaces the tuned fusion
grouping, and "GROUPS": {"<kernel id>": {...}} for per-group options.  Prints the
median per variant and per group kernel.  ``--prebuild`` only compiles the candidates (in parallel, no GPU
needed) so that the variant cache travels to the GPU box with the snapshot.
"""
import json
import os
import statistics
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DOMAIN = (1024, 1024, 80)


def candidate(kernel, extra):
    """The tuned program of ``kernel`` with the option set ``extra`` applied."""
    from absinthe_b200 import programs as P
    cfg = json.load(open(os.path.join(ROOT, "absinthe_b200", "tuned.json")))[kernel]
    extra = dict(extra)
    groups = P.split_sequence(P.KERNELS[kernel][1](), extra.pop("_assignment", cfg["assignment"]))
    tiles = [tuple(t) for t in cfg["tiles"]]
    shapes = extra.pop("_shapes", None)
    if shapes:
        shapes = shapes * len(groups) if len(shapes) == 1 else shapes
        tiles = [tuple(-(-d // s) for d, s in zip(DOMAIN, shape)) for shape in shapes]
    program = P.make_program(kernel, groups, tiles, domain=DOMAIN)
    program["CUDA"] = {} if extra.pop("_clean", False) else dict(cfg["cuda"])
    program["CUDA"].update(extra)
    return program


def prebuild(job):
    from absinthe_b200.generator import build_variant
    try:
        return build_variant(candidate(*job))[1]
    except Exception as exc:
        return "FAILED %s: %s" % (json.dumps(job[1]), str(exc)[-300:])


def main():
    args = sys.argv[1:]
    build_only = "--prebuild" in args
    check = "--check" in args
    args = [a for a in args if a not in ("--prebuild", "--check")]
    texts = []
    for a in args[1:]:                      # "@file": one JSON option set per line
        texts += [line for line in open(a[1:]).read().splitlines() if line.strip()] if a.startswith("@") else [a]
    kernel, overrides = args[0], [json.loads(t) for t in texts]
    if build_only:
        with ProcessPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
            for line in pool.map(prebuild, [(kernel, extra) for extra in overrides]):
                print(line)
        return
    import torch
    from absinthe_b200.launcher import Variant, make_periodic
    variants, tensors = [], {}
    gen = torch.Generator(device="cuda").manual_seed(7)
    for extra in overrides:
        print("loading", json.dumps(extra), flush=True)
        v = Variant.from_program(candidate(kernel, extra))
        dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
        for n in v.inputs + v.outputs:
            if n not in tensors:
                t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
                make_periodic(t, dims)
                tensors[n] = t
        v.bind([tensors[n] for n in v.inputs], [tensors[n] for n in v.outputs])
        variants.append(v)
    if check:
        # every candidate must produce the same bits as the first one (inputs and outputs are shared tensors)
        want = None
        for extra, v in zip(overrides, variants):
            for n in v.outputs:
                tensors[n].zero_()
            v.apply()
            torch.cuda.synchronize()
            got = {n: tensors[n].clone() for n in v.outputs}
            if want is None:
                want = got
            else:
                bad = [n for n in want if not torch.equal(want[n], got[n])]
                print("CHECK %s %s" % ("differs in " + ",".join(bad) if bad else "bit-identical", json.dumps(extra)),
                      flush=True)
        del want, got
    samples = [[] for _ in variants]
    per_group = [[[] for _ in range(v.n_groups)] for v in variants]
    for rnd in range(8):
        for n, v in enumerate(variants):
            total, _ = v.run(6, flush=False)
            events = []
            for _ in range(4):
                for g in range(1, v.n_groups + 1):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    v.apply_group(g)
                    b.record()
                    events.append((g, a, b))
            torch.cuda.synchronize()
            if rnd >= 2:
                samples[n] += total
                for g, a, b in events:
                    per_group[n][g - 1].append(a.elapsed_time(b))
    for extra, s, groups_ms in zip(overrides, samples, per_group):
        print("%-60s median %.4f ms  min %.4f  groups %s" % (
            json.dumps(extra), statistics.median(s), min(s),
            " ".join("%.4f" % statistics.median(g) for g in groups_ms)), flush=True)


if __name__ == "__main__":
    main()
