#!/usr/bin/env python
"""Time the tuned variants of absinthe_b200/tuned.json: short (4 runs) and sustained (60 runs) medians."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from absinthe_b200 import programs as P
from absinthe_b200.analysis import algorithmic_bytes_per_point, analyse
from absinthe_b200.launcher import Variant, make_periodic

def main():
    tuned = json.load(open(os.path.join(ROOT, "absinthe_b200", "tuned.json")))
    domain = (1024, 1024, 80)
    gen = torch.Generator(device="cuda").manual_seed(7)
    for kernel in sys.argv[1:] or sorted(tuned):
        cfg = tuned[kernel]
        seq = P.KERNELS[kernel][1]()
        program = P.make_program(kernel, P.split_sequence(seq, cfg["assignment"]), [tuple(t) for t in cfg["tiles"]], domain=domain)
        program["CUDA"] = dict(cfg.get("cuda", {}))
        v = Variant.from_program(P.clone(program))
        dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
        ts = {}
        for n in v.inputs + v.outputs:
            t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
            make_periodic(t, dims); ts[n] = t
        v.bind([ts[n] for n in v.inputs], [ts[n] for n in v.outputs])
        bpp = algorithmic_bytes_per_point(analyse(P.clone(program)))
        pts = domain[0] * domain[1] * domain[2]
        short = statistics.median(v.run(4, flush=False)[0])
        long_ = statistics.median(v.run(60, flush=False)[0][30:])
        print("%-10s short %.4f ms (%.0f GB/s)   sustained %.4f ms (%.0f GB/s)" % (
            kernel, short, bpp * pts / short / 1e6, long_, bpp * pts / long_ / 1e6), flush=True)
        v.close(); del ts
        torch.cuda.empty_cache()
if __name__ == "__main__":
    main()
