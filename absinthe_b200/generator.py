"""The generator surface of the reference, retargeted at sm_100a.

Same entry points and argument meaning as ``stencil_generator.py`` of the
reference -- ``generate_code(template, name, program)`` (ref :41-59),
``generate_makefile`` (:62-89), ``build_experiment`` (:92-94),
``generate_script`` (:97-109), ``parse_results`` (:112-158), ``write_results``
(:161-177) -- so the experiment drivers keep their ``-g/-b/-p/-f`` recipe:

* ``generate_code("template_tiling.cu" | "template_training.cu", folder + name + ".cu", program)``
  analyses the program *in place* (like the reference), renders the CUDA source and writes a
  launcher descriptor ``<name>.desc`` beside it.
* the Makefile compiles every ``<name>.cu`` to ``<name>.cubin`` with
  ``nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false``.
* ``run.sh`` executes ``python -m absinthe_b200.runner <name>`` per experiment; the runner
  prints the reference's stdout protocol, so ``parse_results`` / ``results.csv`` are unchanged.
"""

import csv
import hashlib
import math
import os
import subprocess
import sys
from collections import deque

from jinja2 import Environment, FileSystemLoader

from copy import deepcopy

from . import analysis as A
from .emitter import Plan

TEMPLATE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "templates")
CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "build", "variants")
CUDA_TEMPLATES = ("template_tiling.cu", "template_training.cu")

TOTAL_PREFIX = "   - total time (min/median/max) [ms]: "
HALO_PREFIX = "   - halo time (min/median/max) [ms]: "
DOMAIN_PREFIX = "   - domain "
VARIANT_PREFIX = "   - variant "


def compute_schedule(program):
    """PUT / WAIT / COMP entries in program order (ref stencil_generator.py:13-38).

    A group's halo exchange is put right after its computation and awaited right before
    the *next* group starts, which is where the multi-GPU driver overlaps it.
    """
    schedule = []
    inflight = deque([None])
    for group in program["TILING"]["GROUPS"]:
        awaited = inflight.popleft()
        if awaited:
            schedule.append(awaited)
        if group["LOOPS"]:
            schedule.append({"TYPE": "COMP", "GROUP": group})
        if group["HALOS"]:
            schedule.append({"TYPE": "PUT", "GROUP": group})
            inflight.append({"TYPE": "WAIT", "GROUP": group})
        else:
            inflight.append(None)
    program["SCHEDULE"] = schedule


def _environment():
    # templates are looked up in the working directory first (the reference's convention,
    # ref stencil_generator.py:54) and then in the package
    return Environment(loader=FileSystemLoader([os.getcwd(), TEMPLATE_DIR]), keep_trailing_newline=True)


def render(template, program):
    """Analyse ``program`` in place and return (cuda source, descriptor text, nvcc flags)."""
    if template not in CUDA_TEMPLATES:
        raise ValueError("absinthe_b200 renders %s, not '%s' (the C++/OpenMP templates belong to the "
                         "reference)" % (" / ".join(CUDA_TEMPLATES), template))
    decompose(program)
    A.analyse(program)
    if "SCHEDULE" not in program:
        compute_schedule(program)
    if template == "template_training.cu":
        program.setdefault("TRAINING", {"STRIDE": 20})
    plan = Plan(program)
    context = plan.context()
    context["TRAINING"] = template == "template_training.cu"
    source = _environment().get_template(template).render(context)
    return source, plan.descriptor_text(), plan.nvcc_flags()


def decompose(program):
    """Apply the node grid ``TILING.NX/NY/NZ`` (ref template_tiling.cpp:189-203): the kernels of one rank work
    on its subdomain.  The backend decomposes along j only (rows stay contiguous, k stays whole for the vertical
    dependences): ``X, Y, Z`` become the slab, ``GLOBAL`` keeps the decomposed extent."""
    tiling = program["TILING"]
    nodes = tuple(tiling.get(k, 1) for k in ("NX", "NY", "NZ"))
    if nodes == (1, 1, 1) or "GLOBAL" in program:
        return
    assert nodes[0] == 1 and nodes[2] == 1, "the B200 backend decomposes along j only (TILING.NX = TILING.NZ = 1)"
    assert program["Y"] % nodes[1] == 0, "Y must be divisible by TILING.NY"
    program["GLOBAL"] = (program["X"], program["Y"], program["Z"])
    program["Y"] //= nodes[1]


REFERENCE_SUFFIX = "_reference"
REFERENCE_TILE = (128, 8, 4)      # points per tile of the sequential reference variant


def sequential_reference(program):
    """The program of ``compute_reference`` (ref template_tiling.cpp:162-177) as a variant of its own.

    Every stencil becomes a group, in the order the reference walks them (groups in TILING order,
    stencils in group order); the launcher applies the periodic update to every group output, i.e.
    after each stencil, and intermediates live in global memory -- the sequential algorithm on the
    GPU.  A ``VERIFY`` program is checked against it at run time (ref :276-282, :378-396).
    """
    order = []
    for outer in program["TILING"]["GROUPS"]:
        for inner in outer.get("GROUPS", []):
            order += [s["NAME"] if isinstance(s, dict) else s for s in inner["STENCILS"]]
    ref = {k: deepcopy(v) for k, v in program.items() if k not in ("TILING", "SCHEDULE", "CUDA")}
    # the plainest kernels the emitter has, so that the check shares as little as possible with the variant it
    # checks: staged __ldg loads instead of TMA, intermediates in global memory, nothing inlined or shuffled,
    # periodic copies by the launcher's own kernel, strict arithmetic
    ref["CUDA"] = {"LOAD": "ldg", "FMAD": False}
    ref["SEQUENCE"] = list(order)
    ref["VARIANT"] = "%s%s" % (program.get("VARIANT") or program.get("NAME") or "variant", REFERENCE_SUFFIX)
    ref["VERIFY"] = False
    counts = tuple(max(1, math.ceil(program[d] / t)) for d, t in zip("XYZ", REFERENCE_TILE))
    tiling = program["TILING"]
    ref["TILING"] = {
        "NX": tiling.get("NX", 1), "NY": tiling.get("NY", 1), "NZ": tiling.get("NZ", 1),
        "GROUPS": [{"GROUPS": [{"STENCILS": [name], "NX": counts[0], "NY": counts[1], "NZ": counts[2]}]}
                   for name in order]}
    return ref


def generate_code(template, name, program):
    """Render ``program`` to ``name`` (a .cu path) and write ``<stem>.desc`` beside it.

    A program with ``VERIFY`` set also gets ``<stem>_reference.cu/.desc`` (``sequential_reference``),
    which the runner executes once and compares the outputs with (ref template_tiling.cpp:378-396).
    """
    stem = name[:-3] if name.endswith(".cu") else name
    if program.get("VERIFY") and template == "template_tiling.cu" and not stem.endswith(REFERENCE_SUFFIX):
        generate_code(template, stem + REFERENCE_SUFFIX + ".cu", sequential_reference(program))
    source, descriptor, flags = render(template, program)
    with open(stem + ".cu", "w") as handle:
        handle.write(source)
    with open(stem + ".desc", "w") as handle:
        handle.write(descriptor)
    program.setdefault("CUDA", {})["NVCC_FLAGS"] = flags


def nvcc_command(flags, source, target):
    return ["nvcc"] + list(flags) + ["-cubin", "-o", target, source]


def generate_makefile(name, experiments):
    """One ``<experiment>.cubin`` rule per experiment (ref stencil_generator.py:62-89).  Flags are per target: an
    experiment that opts into FMA contraction (``CUDA.FMAD``) must not cost the others their ``-fmad=false``."""
    base = ["-std=c++17", "-O3", "-lineinfo", "-gencode arch=compute_100a,code=sm_100a"]
    with open(name, "w", newline="\n") as handle:
        handle.write("\nNVCC=nvcc\n")
        handle.write("NVCCFLAGS= \\\n\t" + " \\\n\t".join(base) + "\n")
        handle.write("# bit-exact with the strict oracle: no contraction of a*b+c into fma\n")
        handle.write("STRICT=-fmad=false\n\n")
        # a VERIFY experiment comes with its sequential reference variant (generate_code)
        keys = [(key + tail, e) for key, e in experiments.items()
                for tail in ([""] + ([REFERENCE_SUFFIX] if e.get("VERIFY") and not e.get("TRAINING") else []))]
        handle.write("all: " + " ".join(key + ".cubin" for key, _ in keys) + "\n\n")
        for key, e in keys:
            strict = not e.get("CUDA", {}).get("FMAD", False) or key.endswith(REFERENCE_SUFFIX)
            handle.write("%s.cubin: %s.cu\n" % (key, key))
            handle.write("\t$(NVCC) $(NVCCFLAGS)%s -cubin -o $@ $<\n\n" % (" $(STRICT)" if strict else ""))
        handle.write("clean:\n\trm -f *.cubin\n")


def build_experiment(name):
    """Build everything in the working directory, like the reference (ref :92-94)."""
    subprocess.call(["make", "-j", str(os.cpu_count() or 1)])


def generate_script(name, experiments, cores, runs):
    """``run.sh``: one runner invocation per experiment and repetition (ref :97-109)."""
    with open(name, "w", newline="\n") as handle:
        handle.write("#!/bin/bash -l\n")
        handle.write("# CORES in the variant description = %s SMs of the target GPU\n\n" % cores)
        package_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        handle.write("export PYTHONPATH=%s:$PYTHONPATH\n\n" % package_root)
        # one interpreter start per repetition: the runner takes any number of variants and prints
        # the same per-variant protocol as one process each would.  Experiments with a node grid
        # (TILING.NY ranks along j) run under torchrun, one rank per GPU.
        ranks = max([e.get("TILING", {}).get("NY", 1) for e in experiments.values()] or [1])
        launch = "%s -m absinthe_b200.runner" % os.path.basename(sys.executable)
        if ranks > 1:
            launch = ("%s -m torch.distributed.run --nnodes=1 --nproc-per-node %d --master-addr 127.0.0.1 "
                      "--master-port 29517 -m absinthe_b200.runner" % (os.path.basename(sys.executable), ranks))
        for _ in range(runs):
            handle.write(launch + " \\\n")
            handle.write(" \\\n".join("    ./" + key for key in experiments) + "\n")


def parse_results(results, skip=16):
    """stdout of a run -> rows ``[VAR, X, Y, Z, TOTAL, HALO]`` (ref stencil_generator.py:112-158).

    A row is produced by every ``halo time`` line once ``skip`` samples of the current
    binary (counted from its ``domain`` line) have been dropped.
    """
    rows, variant, domain, total, remaining = [], None, [], [], 0
    with open(results, "r") as handle:
        for line in handle:
            if line.startswith(DOMAIN_PREFIX):
                domain = [int(v) for v in line[len(DOMAIN_PREFIX):].split(", ")]
                remaining = skip
            elif line.startswith(VARIANT_PREFIX):
                variant = line[len(VARIANT_PREFIX):].rstrip()
            elif line.startswith(TOTAL_PREFIX):
                total = [float(v) for v in line[len(TOTAL_PREFIX):].split("/")]
            elif line.startswith(HALO_PREFIX):
                halo = [float(v) for v in line[len(HALO_PREFIX):].split("/")]
                if remaining == 0:
                    rows.append([variant] + [str(v) for v in domain] + [str(v) for v in total]
                                + [str(v) for v in halo])
                remaining = max(0, remaining - 1)
    return rows


def write_results(rows, table):
    """rows -> ``VAR,X,Y,Z,TOTAL,HALO`` csv (ref stencil_generator.py:161-177)."""
    print("-> writing results to " + table)
    with open(table, "w") as handle:
        out = csv.writer(handle, delimiter=",", quotechar="'", lineterminator="\n")
        out.writerow(["VAR", "X", "Y", "Z", "TOTAL", "HALO"])
        for row in rows:
            out.writerow(row)


# ---------------------------------------------------------------------------------------
# In-process build with a content-addressed cache (tests, bench, auto-tuner).
# ---------------------------------------------------------------------------------------
def build_variant(program, template="template_tiling.cu", cache_dir=None):
    """Render + compile ``program``; returns (descriptor path, cubin path).  Cached by content."""
    source, descriptor, flags = render(template, program)
    tag = hashlib.sha256((source + descriptor + " ".join(flags)).encode()).hexdigest()[:24]
    cache_dir = os.path.abspath(cache_dir or CACHE_DIR)
    os.makedirs(cache_dir, exist_ok=True)
    stem = os.path.join(cache_dir, tag)
    if not (os.path.exists(stem + ".cubin") and os.path.exists(stem + ".desc")):
        # several ranks may build the same variant at once: everything goes through per-process
        # temporaries and is published with atomic renames.  The source is published first and compiled
        # under its final name (the tag is its content hash, so concurrent writers agree byte for byte):
        # -lineinfo then points at a file that exists, which `ncu --import-source on` needs.
        mine = "%s.%d" % (stem, os.getpid())
        with open(mine + ".cu", "w") as handle:
            handle.write(source)
        os.replace(mine + ".cu", stem + ".cu")
        proc = subprocess.run(nvcc_command(flags, stem + ".cu", mine + ".tmp"), capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (stem + ".cu", proc.stderr[-4000:]))
        with open(mine + ".desc.tmp", "w") as handle:
            handle.write(descriptor)
        os.replace(mine + ".tmp", stem + ".cubin")
        os.replace(mine + ".desc.tmp", stem + ".desc")
    return stem + ".desc", stem + ".cubin"
