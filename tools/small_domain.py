#!/usr/bin/env python
"""Latency of whole sequences on the reference's default domain (64 x 64 x 60): stream launches vs one
CUDA graph per sequence.  The domain fits L2 many times over; this is a launch-latency measurement."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from absinthe_b200 import programs as P
from absinthe_b200.launcher import Variant, make_periodic

def main():
    d, a, f = (P.KERNELS[k][1]() for k in ("diffusion", "advection", "fastwaves"))
    fast = {"INLINE": "all", "FUSE_HALO": True, "BLOCK": (2, 1), "SHUFFLE": True, "THREADS": 256}
    cases = [
        ("diffusion fused", P.make_program("diffusion", [d], [(2, 16, 5)]), fast),
        ("diffusion 4 groups", P.make_program("diffusion", P.split_sequence(d, [0]*4+[1]*4+[2]*4+[3]*4), [(2, 16, 5)]*4), fast),
        ("diffusion unfused (16 groups)", P.make_program("diffusion", [[s] for s in d], [(2, 16, 5)]*16), fast),
        ("advection fused", P.make_program("advection", [a], [(2, 16, 5)]), dict(fast, SHUFFLE=False)),
        ("fastwaves 2 groups", P.make_program("fastwaves", [f[:6], f[6:]], [(2, 16, 5)]*2), dict(fast, SHUFFLE=False, BLOCK=(1, 2))),
    ]
    out = []
    gen = torch.Generator(device="cuda").manual_seed(1)
    for name, program, cuda in cases:
        program["CUDA"] = dict(cuda)
        v = Variant.from_program(program)
        dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
        ins = []
        for _ in v.inputs:
            t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
            make_periodic(t, dims); ins.append(t)
        outs = [v.empty_field() for _ in v.outputs]
        v.bind(ins, outs)
        row = {"case": name, "launches": v.launches}
        for graph in (False, True):
            for _ in range(20): v.apply(graph=graph)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a_.record()
            for _ in range(500): v.apply(graph=graph)
            b_.record(); torch.cuda.synchronize()
            us = a_.elapsed_time(b_) * 1e3 / 500
            row["graph_us" if graph else "stream_us"] = round(us, 2)
        row["mpoints_s_graph"] = round(v.points / row["graph_us"], 1)
        out.append(row); print(json.dumps(row), flush=True)
        v.close()
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
if __name__ == "__main__":
    main()
