"""Worker of the two-ranks-on-one-GPU exchange test: each process owns a j-slab, handles travel through gloo."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def program_for(kernel, slab):
    from absinthe_b200 import programs as P
    seq = P.KERNELS[kernel][1]()
    if kernel == "fastwaves":
        program = P.make_program(kernel, [seq[:6], seq[6:]], [(2, 3, 2), (1, 2, 5)], domain=slab)
        program["CUDA"] = {"INLINE": "all", "FUSE_HALO": True, "BLOCK": (1, 2),
                           "GROUPS": {"2": {"STREAM_K": "regs", "BLOCK": (2, 1), "COLUMN": {"WARPS": (1, 2)}}}}
    else:
        program = P.make_program(kernel, [seq[:8], seq[8:]], [(2, 2, 3), (1, 4, 2)], domain=slab)
        program["CUDA"] = {"FUSE_HALO": False}
    return program


def worker(rank, world, port, kernel, slab, steps, result, engine=None):
    import torch
    import torch.distributed as dist
    from absinthe_b200 import programs as P
    from absinthe_b200.distributed import P2PExchange
    from absinthe_b200.launcher import Variant
    from seam_check import field_seed, global_field
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        device = rank % torch.cuda.device_count()
        torch.cuda.set_device(device)
        v = Variant.from_program(P.clone(program_for(kernel, slab)), device=device)
        X, Y, Z = slab
        extent = (X, world * Y, Z)
        mine = ((-3, X + 3), (rank * Y - 3, rank * Y + Y + 3), (-3, Z + 3))
        ins = [global_field(field_seed(kernel, n), mine, extent, "cuda:%d" % device) for n in v.inputs]
        outs = [v.empty_field() for _ in v.outputs]
        v.bind(ins, outs)
        exchange = P2PExchange(v, world, rank)
        for _ in range(steps):                      # repeated steps exercise both parities of the windows
            for g in range(1, v.n_groups + 1):
                exchange.wait(g)
                if g > 1:
                    exchange.wait(g - 1)
                exchange.apply(g, engine)         # None: the defaults (fused where the successor reads the halos)
        for g in range(1, v.n_groups + 1):
            exchange.wait(g)
        torch.cuda.synchronize()
        np.savez(result % rank, **{n: t.cpu().numpy() for n, t in zip(v.outputs, outs)})
        dist.barrier()
        v.close()
    finally:
        dist.destroy_process_group()
