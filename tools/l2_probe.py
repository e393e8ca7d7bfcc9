#!/usr/bin/env python
"""Does the L2 keep what fast-waves group 1 wrote until group 2 reads it?  Times both group kernels of the tuned
two-group variant on domains 1024 x Y x 80 of growing Y (same tile shapes): if group 2 finds uout / vout / wgtfac in
L2, its time per point drops below the full-domain figure.

    l2_probe.py [--prebuild] Y [Y ...]
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = ((256, 4, 40), (64, 4, 80))
if os.environ.get("PROBE_SHAPES"):          # e.g. PROBE_SHAPES="64,4,10;64,4,20"
    SHAPES = tuple(tuple(int(x) for x in part.split(",")) for part in os.environ["PROBE_SHAPES"].split(";"))


def program(y):
    from absinthe_b200 import programs as P
    cfg = json.load(open(os.path.join(ROOT, "absinthe_b200", "tuned.json")))["fastwaves"]
    domain = (1024, y, 80)
    groups = P.split_sequence(P.KERNELS["fastwaves"][1](), cfg["assignment"])
    tiles = [tuple(-(-d // s) for d, s in zip(domain, shape)) for shape in SHAPES]
    p = P.make_program("fastwaves", groups, tiles, domain=domain)
    p["CUDA"] = dict(cfg["cuda"])
    return p


def main():
    args = [a for a in sys.argv[1:] if a != "--prebuild"]
    ys = [int(a) for a in args]
    if "--prebuild" in sys.argv:
        from absinthe_b200.generator import build_variant
        for y in ys:
            print(build_variant(program(y))[1])
        return
    import torch
    from absinthe_b200.launcher import Variant, make_periodic
    for y in ys:
        v = Variant.from_program(program(y))
        gen = torch.Generator(device="cuda").manual_seed(7)
        tensors = {}
        for n in v.inputs + v.outputs:
            t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
            make_periodic(t, (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ))
            tensors[n] = t
        v.bind([tensors[n] for n in v.inputs], [tensors[n] for n in v.outputs])
        times = [[], []]
        for rnd in range(30):
            events = []
            for g in (1, 2):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                v.apply_group(g)
                b.record()
                events.append((a, b))
            torch.cuda.synchronize()
            if rnd >= 5:
                for g, (a, b) in enumerate(events):
                    times[g].append(a.elapsed_time(b))
        pts = 1024 * y * 80
        med = [statistics.median(t) for t in times]
        print("Y %5d  field %6.1f MB  g1 %.4f ms = %.3f ns/pt (%.0f GB/s on 80 B/pt)   g2 %.4f ms = %.3f ns/pt (%.0f GB/s on 48 B/pt)"
              % (y, 1030 * (y + 6) * 86 * 8 / 1e6, med[0], med[0] * 1e6 / pts, 80 * pts / med[0] / 1e6, med[1],
                 med[1] * 1e6 / pts, 48 * pts / med[1] / 1e6), flush=True)
        v.close()
        del tensors


if __name__ == "__main__":
    main()
