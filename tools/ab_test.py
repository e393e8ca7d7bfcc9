#!/usr/bin/env python
"""Interleaved A/B timing of CUDA option sets on one tuned variant (removes warm-up / power-state order bias).
usage: ab_test.py <kernel> '<json options A>' '<json options B>' ..."""
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
    for extra in overrides:
        program = P.make_program(kernel, P.split_sequence(seq, cfg["assignment"]), [tuple(t) for t in cfg["tiles"]], domain=(1024, 1024, 80))
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
    for rnd in range(8):
        for n, v in enumerate(variants):
            total, _ = v.run(6, flush=False)
            if rnd >= 2:
                samples[n] += total
    for extra, s in zip(overrides, samples):
        print("%-60s median %.4f ms  min %.4f" % (json.dumps(extra), statistics.median(s), min(s)), flush=True)

if __name__ == "__main__":
    main()
