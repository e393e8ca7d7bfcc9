#!/usr/bin/env python
"""Interleaved timing of prebuilt variants given as (descriptor, cubin) paths: build/ab_paths.json =
{kernel: {label: [desc, cubin]}} -- for comparing template revisions (tools/ab_test.py compares option sets)."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from absinthe_b200.launcher import Variant, make_periodic
    table = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "build", "ab_paths.json")))
    for kernel, entries in table.items():
        gen = torch.Generator(device="cuda").manual_seed(7)
        tensors, variants = {}, []
        for label, (desc, cubin) in entries.items():
            v = Variant(os.path.join(ROOT, os.path.relpath(desc, ROOT)) if os.path.isabs(desc) and not os.path.exists(desc) else desc, cubin)
            dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
            for n in v.inputs + v.outputs:
                if n not in tensors:
                    t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
                    make_periodic(t, dims)
                    tensors[n] = t
            v.bind([tensors[n] for n in v.inputs], [tensors[n] for n in v.outputs])
            variants.append((label, v))
        samples = {label: [] for label, _ in variants}
        for rnd in range(10):
            for label, v in variants:
                total, _ = v.run(6, flush=False)
                if rnd >= 2:
                    samples[label] += total
        for label, _ in variants:
            print("%-10s %-6s median %.4f ms  min %.4f" % (kernel, label, statistics.median(samples[label]), min(samples[label])), flush=True)
        for _, v in variants:
            v.close()
        del tensors
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
