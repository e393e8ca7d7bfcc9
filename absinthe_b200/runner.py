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


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("stem", nargs="+", help="path(s) of the variant(s) without extension")
    parser.add_argument("--runs", type=int, default=None, help="timed runs (default: RUNS of the program)")
    parser.add_argument("--no-flush", action="store_true")
    parser.add_argument("--device", type=int, default=0)
    args = parser.parse_args(argv)
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
        run_variant(stem, runs, flush=flush, training=training, device=args.device,
                    verify=meta.get("verify", "0") == "1")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
