"""torchrun entry: N ranks, each a j-slab, must reproduce the single-GPU result of the global domain."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from absinthe_b200 import programs as P  # noqa: E402
from absinthe_b200.distributed import SlabExchange  # noqa: E402
from absinthe_b200.launcher import Variant, make_periodic  # noqa: E402


def run_case(kernel, groups, tiles, local_domain, cuda, rank, world):
    X, Y, Z = local_domain
    seq = P.KERNELS[kernel][1]()
    local_prog = P.make_program(kernel, groups, tiles, domain=(X, Y, Z))
    global_prog = P.make_program(kernel, groups, tiles, domain=(X, world * Y, Z))
    local_prog["CUDA"] = dict(cuda)
    global_prog["CUDA"] = dict(cuda)
    vg = Variant.from_program(global_prog, device=rank)
    vl = Variant.from_program(local_prog, device=rank)
    gen = torch.Generator(device="cuda").manual_seed(99)
    gdims = (vg.X, vg.Y, vg.Z, vg.HX, vg.HY, vg.HZ)
    gin = []
    for _ in vg.inputs:
        t = torch.rand(vg.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
        make_periodic(t, gdims)
        gin.append(t)
    gout = [vg.empty_field() for _ in vg.outputs]
    vg.bind(gin, gout)
    vg.apply()
    # the rank's slab of the (periodic) global inputs, halos included
    lin = [t[:, rank * Y:rank * Y + Y + 6, :].contiguous() for t in gin]
    lout = [vl.empty_field() for _ in vl.outputs]
    vl.bind(lin, lout)
    exchange = SlabExchange(vl, world, rank)
    for g in range(1, vl.n_groups + 1):
        if g > 1:
            exchange.wait(g - 1)
        vl.apply_group(g, with_halo=False)
        if g % 2:
            exchange.update_async(g)          # both flavours of the PUT/WAIT pair are exercised
        else:
            exchange.update(g)
    for g in range(1, vl.n_groups + 1):
        exchange.wait(g)
    torch.cuda.synchronize()
    for name, mine, whole in zip(vl.outputs, lout, gout):
        want = whole[:, rank * Y:rank * Y + Y + 6, :]
        assert torch.equal(mine, want), "%s rank %d: %s differs (%d cells)" % (
            kernel, rank, name, int((mine != want).sum()))
    vl.close()
    vg.close()


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(rank)
    d, f = P.KERNELS["diffusion"][1](), P.KERNELS["fastwaves"][1]()
    fast = {"INLINE": "all", "BLOCK": (2, 1), "SHUFFLE": True, "FUSE_HALO": True}
    run_case("diffusion", [d], [(2, 3, 4)], (64, 24, 12), {}, rank, world)
    run_case("diffusion", [d], [(2, 3, 4)], (64, 24, 12), fast, rank, world)
    run_case("fastwaves", [f[:6], f[6:]], [(1, 4, 2), (2, 2, 3)], (48, 20, 10), {}, rank, world)
    run_case("fastwaves", [f[:6], f[6:]], [(1, 4, 2), (2, 2, 3)], (48, 20, 10), fast, rank, world)
    run_case("advection", [P.KERNELS["advection"][1]()], [(1, 2, 2)], (40, 16, 6), fast, rank, world)
    dist.barrier()
    if rank == 0:
        print("multi-gpu parity ok on %d ranks" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
