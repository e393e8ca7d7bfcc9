#!/usr/bin/env python
"""Apply one A/B candidate (tools/ab_test.py syntax) a few times -- the command ncu wraps.

    profile_candidate.py <kernel> '<json options>' [applications]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    from ab_test import candidate
    from absinthe_b200.launcher import Variant, make_periodic
    kernel, extra = sys.argv[1], json.loads(sys.argv[2])
    times = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    v = Variant.from_program(candidate(kernel, extra))
    dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
    gen = torch.Generator(device="cuda").manual_seed(7)
    tensors = {}
    for n in v.inputs + v.outputs:
        t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
        make_periodic(t, dims)
        tensors[n] = t
    v.bind([tensors[n] for n in v.inputs], [tensors[n] for n in v.outputs])
    for _ in range(times):
        v.apply()
    torch.cuda.synchronize()
    print("applied", times, "times")


if __name__ == "__main__":
    main()
