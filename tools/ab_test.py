#!/usr/bin/env python
"""Interleaved A/B timing of CUDA option sets on one tuned variant (removes warm-up / power-state order bias).
usage: ab_test.py <kernel> '<json options A>' '<json options B>' ...
An option set may hold "_shapes": [[x, y, z], ...] -- tile shapes in points, one per group (or one for all) --
which replaces the tuned tile counts.  Prints the median per variant and per group."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from absinthe_b200 import programs as P
from absinthe_b200.launcher import Variant, make_periodic

def main():
    kernel = sys.argv[1]
    overrides = [json.loads(a) for a in sys.argv[2:]]
    cfg = json.load(open(os.path.join(ROOT, "absinthe_b200", "tuned.json")))[kernel]
    seq = P.KERNELS[kernel][1]()
    variants, tensors = [], {}
    gen = torch.Generator(device="cuda").manual_seed(7)
    domain = (1024, 1024, 80)
    groups = P.split_sequence(seq, cfg["assignment"])
    for extra in overrides:
        extra = dict(extra)
        tiles = [tuple(t) for t in cfg["tiles"]]
        shapes = extra.pop("_shapes", None)
        if shapes:
            shapes = shapes * len(groups) if len(shapes) == 1 else shapes
            tiles = [tuple(-(-d // s) for d, s in zip(domain, shape)) for shape in shapes]
        program = P.make_program(kernel, groups, tiles, domain=domain)
        program["CUDA"] = dict(cfg["cuda"]); program["CUDA"].update(extra)
        v = Variant.from_program(program)
        dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
        for n in v.inputs + v.outputs:
            if n not in tensors:
                t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
                make_periodic(t, dims); tensors[n] = t
        v.bind([tensors[n] for n in v.inputs], [tensors[n] for n in v.outputs])
        variants.append(v)
    samples = [[] for _ in variants]
    per_group = [[[] for _ in range(v.n_groups)] for v in variants]
    for rnd in range(8):
        for n, v in enumerate(variants):
            total, _ = v.run(6, flush=False)
            events = []
            for _ in range(4):
                for g in range(1, v.n_groups + 1):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); v.apply_group(g); b.record()
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
