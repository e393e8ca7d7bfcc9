"""Run one generated variant and print the reference's stdout protocol.

``python -m absinthe_b200.runner <stem>`` loads ``<stem>.desc`` + ``<stem>.cubin`` and plays the role
of the reference's generated binary (ref template_tiling.cpp:179-397): inputs from
``srand(0)``/``rand()`` made periodic, outputs zeroed, ``2*RUNS`` applications with an optional L2
flush before each, one ``total time`` / ``halo time`` line pair for every odd run.  The lines are
byte-compatible with what ``parse_results`` expects (ref stencil_generator.py:127-138).
"""

import argparse
import sys

import numpy as np


def prepare_inputs(variant, seed=0, on_device=False):
    """Device tensors initialised exactly like the reference's main() (ref :209-221).

    ``on_device`` draws the values on the GPU instead (training sweeps: the domains are large, the
    values irrelevant, and a host-side rand() fill of up to 27 fields dominated the run time)."""
    import torch
    from .launcher import make_periodic, reference_inputs
    dims = (variant.X, variant.Y, variant.Z, variant.HX, variant.HY, variant.HZ)
    if on_device:
        gen = torch.Generator(device="cuda:%d" % variant.device).manual_seed(seed)
        tensors = []
        for _ in variant.inputs:
            t = torch.rand(variant.shape, dtype=torch.float64, device="cuda:%d" % variant.device, generator=gen)
            make_periodic(t, dims)
            tensors.append(t)
        return tensors
    host = reference_inputs(len(variant.inputs), variant.shape, seed)
    tensors = []
    for array in host:
        t = torch.from_numpy(array).to("cuda:%d" % variant.device)
        make_periodic(t, dims)
        tensors.append(t)
    return tensors


def compute_reference(stem, inputs, device=0):
    """Outputs of the sequential reference variant ``<stem>_reference`` (generator.sequential_reference:
    every stencil over the whole domain, periodic update after each -- ref template_tiling.cpp:162-177)."""
    from .generator import REFERENCE_SUFFIX
    from .launcher import Variant
    reference = Variant(stem + REFERENCE_SUFFIX + ".desc", stem + REFERENCE_SUFFIX + ".cubin", device)
    expected = [reference.empty_field() for _ in reference.outputs]
    reference.bind(inputs, expected)
    reference.apply()
    names = list(reference.outputs)
    reference.close()
    return dict(zip(names, expected))


def run_variant(stem, runs, flush=True, training=False, device=0, out=sys.stdout, verify=False):
    import torch
    from .launcher import Variant, compare_fields
    variant = Variant(stem + ".desc", stem + ".cubin", device)
    name = open(stem + ".desc").read().split("\n")[1].split(" ", 1)[1]
    print("-> configuration", file=out)
    print("   - variant " + name, file=out)
    print("   - domain %d, %d, %d" % (variant.X, variant.Y, variant.Z), file=out)
    print("   - runs %d" % runs, file=out)
    print("   - verify %s" % bool(verify), file=out)
    print("   - device " + torch.cuda.get_device_name(device), file=out)
    inputs = prepare_inputs(variant, on_device=training)
    outputs = [variant.empty_field() for _ in variant.outputs]
    expected = None
    if verify:
        print("-> computing reference...", file=out)
        expected = compute_reference(stem, dict(zip(variant.inputs, inputs)), device)
    variant.bind(inputs, outputs)
    total, halo = variant.run(runs, flush=flush, training=training)
    for t, h in zip(total, halo):
        print("-> apply stencils...", file=out)
        print("   - total time (min/median/max) [ms]: %g" % t, file=out)
        print("   - halo time (min/median/max) [ms]: %g" % h, file=out)
    if verify:
        # ref template_tiling.cpp:378-396: interior of every output, relative 1e-7
        print("-> verifying...", file=out)
        dims = (variant.X, variant.Y, variant.Z, variant.HX, variant.HY, variant.HZ)
        errors = matches = 0
        for field, got in zip(variant.outputs, outputs):
            e, m = compare_fields(expected[field], got, dims)
            errors, matches = errors + e, matches + m
        print("   - %d errors" % errors, file=out)
        print("   - %d matches" % matches, file=out)
    variant.close()
    return np.array(total), np.array(halo)


def run_variant_ranks(stem, runs, rank, world, flush=True, device=0, out=sys.stdout):
    """The same protocol with the domain cut into ``world`` j-slabs, one rank (process) per GPU: the node grid of
    the variant description (``# nodes`` in the descriptor; ref template_tiling.cpp:189-203) filled in.  After
    every group the rank PUTs its edge rows into the neighbours' exchange windows (P2P over NVLink) and WAITs for
    theirs before the next group (ref stencil_generator.py:23-37).  ``total`` is the slowest rank's time per run,
    ``halo`` the part of it not spent in the group kernels.  Rank 0 prints."""
    import torch
    import torch.distributed as dist
    from .distributed import P2PExchange
    from .fields import slab_field
    from .launcher import Variant, flush_l2
    variant = Variant(stem + ".desc", stem + ".cubin", device)
    name = open(stem + ".desc").read().split("\n")[1].split(" ", 1)[1]
    say = (lambda *a: print(*a, file=out)) if rank == 0 else (lambda *a: None)
    say("-> configuration")
    say("   - variant " + name)
    say("   - domain %d, %d, %d" % (variant.X, variant.Y * world, variant.Z))
    say("   - runs %d" % runs)
    say("   - ranks %d (j-slabs of %d rows, P2P halo exchange)" % (world, variant.Y))
    slab, halo = (variant.X, variant.Y, variant.Z), (variant.HX, variant.HY, variant.HZ)
    inputs = [slab_field(name, n, slab, halo, rank, world, "cuda:%d" % device) for n in variant.inputs]
    outputs = [variant.empty_field() for _ in variant.outputs]
    variant.bind(inputs, outputs)
    exchange = P2PExchange(variant, world, rank)
    groups = range(1, variant.n_groups + 1)
    for run in range(2 * runs):
        if flush:
            flush_l2()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(2 + 2 * variant.n_groups)]
        marks[0].record()
        for g in groups:
            if g > 1:
                exchange.wait(g - 1)                 # WAIT: this group reads the halos of the previous one
            marks[2 * g - 1].record()
            exchange.apply(g)                        # COMP + PUT (fused into the kernel, or overlapped with the next one)
            marks[2 * g].record()
        for g in groups:
            exchange.wait(g)
        marks[-1].record()
        torch.cuda.synchronize(device)
        if run % 2 == 1:
            total = marks[0].elapsed_time(marks[-1])
            kernels = sum(marks[2 * g - 1].elapsed_time(marks[2 * g]) for g in groups)
            times = torch.tensor([total, total - kernels], dtype=torch.float64)
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            say("-> apply stencils...")
            say("   - total time (min/median/max) [ms]: %g" % times[0].item())
            say("   - halo time (min/median/max) [ms]: %g" % times[1].item())
    dist.barrier()
    variant.close()
    return outputs


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("stem", nargs="+", help="path(s) of the variant(s) without extension")
    parser.add_argument("--runs", type=int, default=None, help="timed runs (default: RUNS of the program)")
    parser.add_argument("--no-flush", action="store_true")
    parser.add_argument("--device", type=int, default=0)
    args = parser.parse_args(argv)
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:                                    # launched by torchrun: one rank per GPU (ranks share GPUs if fewer)
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo")              # host-side channel for the IPC handles and the timing maxima
        args.device = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
        torch.cuda.set_device(args.device)
    for stem in args.stem:
        meta = {}
        with open(stem + ".desc") as handle:
            for line in handle:
                if line.startswith("# "):
                    key, _, value = line[2:].strip().partition(" ")
                    meta[key] = value
        runs = args.runs or int(meta.get("runs", 64))
        flush = not args.no_flush and meta.get("flush", "1") == "1"
        training = meta.get("training", "0") == "1"
        nodes = [int(t) for t in meta.get("nodes", "1 1 1").split()]
        if world > 1:
            if nodes[1] != world:
                raise SystemExit("%s is generated for TILING.NY = %d ranks, launched with %d" % (stem, nodes[1], world))
            run_variant_ranks(stem, runs, rank, world, flush=flush, device=args.device)
        else:
            if nodes[1] != 1:
                raise SystemExit("%s is decomposed into %d j-slabs: launch it with torchrun --nproc-per-node %d "
                                 "(run.sh does)" % (stem, nodes[1], nodes[1]))
            run_variant(stem, runs, flush=flush, training=training, device=args.device,
                        verify=meta.get("verify", "0") == "1")
        sys.stdout.flush()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
