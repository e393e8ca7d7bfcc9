"""Experiment drivers: the reference's command-line recipe on the B200 backend.

One implementation serves the five scripts at the repository root (``diffusion.py``,
``advection.py``, ``fastwaves.py``, ``fitcache.py``, ``fitddr.py``); flags and file artefacts are the
reference's (ref diffusion.py:298-360, fitcache.py:88-187):

    -a / --auto       brute-force tile-count variants per canonical fusion group   (ref :102-134)
    -o / --optimize   one-shot optimum of the analytical model                     (ref :89-99)
    -e / --explore    OPT / HAND / AUTO / MAX / MIN + neighbouring variants        (ref :137-283)
    -g / --generate   write <folder>/<name>.cu + .desc, Makefile, run.sh
    -b / --build      make
    -p / --parse F    <folder>/F (stdout of run.sh) -> results.csv / auto.csv
    -f / --folder D   working directory

``-o`` and ``-e`` evaluate the analytical model (``optimizer.py``: the reference's MILP, built in process and
solved with HiGHS; ``ABSINTHE_SOLVER=cplex`` shells out to a ``cplex`` like the reference).  An exploration
solves ~175 independent MILPs; they run in a process pool (``ABSINTHE_WORKERS``), each bounded by
``ABSINTHE_MILP_SECONDS``; ``ABSINTHE_SEED`` makes the random sample of groupings reproducible.
"""

import csv
import os
import sys
from copy import deepcopy
from getopt import GetoptError, getopt

from . import programs as P
from . import stencils as S
from .generator import (build_experiment, generate_code, generate_makefile, generate_script, parse_results,
                        write_results)

# canonical fusion groups of the auto-tuning sweeps (ref diffusion.py:106-108, advection.py:96-97,
# fastwaves.py:88-89) and the tile counts they try (ref diffusion.py:114-116)
AUTO_GROUPS = {
    "diffusion": ([["ulap", "ufli", "uflj", "uout"], ["vlap", "vfli", "vflj", "vout"],
                   ["wlap", "wfli", "wflj", "wout"], ["pplap", "ppfli", "ppflj", "ppout"]],
                  [["uout"], ["vout"], ["wout"], ["ppout"]]),
    "advection": ([S.advection_sequence()], [["utev", "vtev"]]),
    "fastwaves": ([["ppgk", "ppgc", "ppgu", "ppgv", "uout", "vout"], ["udc", "vdc", "div"]],
                  [["uout", "vout"], ["div"]]),
}
AUTO_COUNTS_XY = (1, 2, 4, 8, 16, 32, 64)
AUTO_COUNTS_Z = (1, 2, 5, 15, 30, 60)
UNFILTERED = ("fastwaves",)        # ref fastwaves.py:100 has the >= CORES filter commented out


def auto_tune(kernel, experiments, domain=None, counts_xy=AUTO_COUNTS_XY, counts_z=AUTO_COUNTS_Z):
    """One program per (group, nx, ny, nz); named ``<kernel>-<group>-<nx>-<ny>-<nz>`` like the reference."""
    groups, outputs = AUTO_GROUPS[kernel]
    cores = P.base_program(kernel)["MACHINE"]["CORES"]
    for idx, names in enumerate(groups):
        for nx in counts_xy:
            for ny in counts_xy:
                for nz in counts_z:
                    if kernel not in UNFILTERED and nx * ny * nz < cores:
                        continue
                    name = "%s-%d-%d-%d-%d" % (kernel, idx, nx, ny, nz)
                    experiments[name] = P.make_program(kernel, [names], [(nx, ny, nz)], domain=domain,
                                                       variant=name, outputs=outputs[idx], RUNS=16)


# HAND / AUTO of the reference drivers as (group index per stencil, tile counts per group)
# (ref diffusion.py:157-189, advection.py:145-177, fastwaves.py:138-183)
SPECIAL = {
    "diffusion": {"HAND": ([0] * 16, [(1, 1, 60)]), "AUTO": ([n // 4 for n in range(16)], [(1, 1, 60)] * 4)},
    "advection": {"HAND": ([0] * 8, [(1, 1, 60)]), "AUTO": ([0] * 8, [(1, 1, 30)])},
    "fastwaves": {"HAND": ([0] * 6 + [1] * 3, [(1, 8, 8)] * 2), "AUTO": ([0] * 6 + [1] * 3, [(1, 4, 5), (1, 2, 5)])},
}
MAX_GROUPS = 5          # groupings with a group index above 4 are not explored (ref diffusion.py:221)
SAMPLES = 20            # random groupings per exploration (ref diffusion.py:223)


def model_program(kernel, variant, groups=None, tiling=None, name=None, domain=None):
    """A driver-level program ready for ``optimize_program``: sequence fixed, optional CONSTRAINTS."""
    program = P.base_program(kernel)
    program["STENCILS"] = P.KERNELS[kernel][0]()
    program["SEQUENCE"] = P.KERNELS[kernel][1]()
    program["VARIANT"] = variant
    if name:
        program["NAME"] = name
    if domain:
        program["X"], program["Y"], program["Z"] = domain
    if groups is not None:
        program["CONSTRAINTS"]["GROUPS"] = list(groups)
    if tiling is not None:
        program["CONSTRAINTS"]["TILING"] = list(tiling)
    return program


def fixed_tile_counts(program, stencils, counts):
    """CONSTRAINTS.TILING entries that pin the tile counts of ``stencils`` to ``counts``: ``(dim, s, -(c+1))``
    is n <= c and ``(dim, s, c-1)`` is n >= c (ref stencil_optimizer.py:587-594); bounds the model has anyway
    (1 <= n <= extent) are not repeated, like in the reference's HAND / AUTO lists."""
    out = []
    for stencil in stencils:
        for dim, count in zip("xyz", counts):
            if count < program[dim.upper()]:
                out.append((dim, stencil, -(count + 1)))
            if count > 1:
                out.append((dim, stencil, count - 1))
    return out


def special_program(kernel, variant, domain=None):
    """HAND / AUTO: groups and tile counts given, capacity and slack constraints lifted (ref diffusion.py:169-172);
    CORES is lowered to the tile count where the pinned tiling has fewer tiles than the machine has SMs."""
    assignment, tiles = SPECIAL[kernel][variant]
    sequence = P.KERNELS[kernel][1]()
    program = model_program(kernel, variant, groups=list(zip(sequence, assignment)), domain=domain)
    tiling = []
    for index, names in enumerate(P.split_sequence(sequence, assignment)):
        tiling += fixed_tile_counts(program, names, tiles[index])
    program["CONSTRAINTS"]["TILING"] = tiling
    program["MACHINE"]["CAPACITY"] = 2 ** 31
    program["MACHINE"]["CORES"] = min([program["MACHINE"]["CORES"]] + [t[0] * t[1] * t[2] for t in tiles])
    program["SLACK"] = {"SIZE": 1.0, "CORES": 1.0}
    return program


def enumerate_groupings(sequence, limit=MAX_GROUPS):
    """Every assignment of consecutive stencils to groups 0, 1, ... with at most ``limit`` groups."""
    out = [[0]]
    for _ in sequence[1:]:
        out = [g + [g[-1] + step] for g in out for step in (0, 1) if g[-1] + step < limit]
    return out


def grouping_of(kernel, variant):
    """Group indexes encoded in a variant name ``<kernel>-g0-g1-...[-m|p{x,y,z}]`` and its suffix (or None)."""
    sequence = P.KERNELS[kernel][1]()
    parts = variant.split("-")
    digits = [int(t) for t in parts[1:1 + len(sequence)]]
    suffix = parts[1 + len(sequence)] if len(parts) > 1 + len(sequence) else None
    return digits, suffix


def neighbour_tiling(program, solved, suffix):
    """The tile-count neighbours of a solved grouping (ref diffusion.py:240-274): ``-m<dim>`` allows at most
    N - 1 tiles along ``dim`` in every group that has more than one, ``-p<dim>`` demands at least N + 1 in every
    group that has fewer tiles than points, N being the group's tile count in the solution."""
    dim, key = suffix[1], "N" + suffix[1].upper()
    out = []
    for outer in solved["TILING"]["GROUPS"]:
        info = outer["GROUPS"][0]
        if suffix[0] == "m" and info[key] > 1:
            out += [(dim, s, -info[key]) for s in info["STENCILS"]]
        if suffix[0] == "p" and info[key] < program[dim.upper()]:
            out += [(dim, s, info[key]) for s in info["STENCILS"]]
    return out


def _solve(job):
    """One MILP in a worker process: (key, folder, program) -> (key, program) with TILING / OBJECTIVE."""
    import contextlib
    import io
    from .optimizer import optimize_program
    key, folder, program = job
    log = io.StringIO()
    try:
        with contextlib.redirect_stdout(log):
            optimize_program(folder + key, program)
    except Exception as exc:                      # an infeasible neighbour is dropped, like a failed cplex run
        program["ERROR"] = str(exc)[:200]
    return key, program


def solve_all(jobs, workers=None):
    """Solve independent programs in parallel (HiGHS is single threaded; an exploration holds ~175 of them)."""
    from concurrent.futures import ProcessPoolExecutor
    workers = workers or int(os.environ.get("ABSINTHE_WORKERS", os.cpu_count() or 1))
    if workers <= 1 or len(jobs) <= 1:
        return [_solve(job) for job in jobs]
    with ProcessPoolExecutor(max_workers=workers) as pool:
        return list(pool.map(_solve, jobs))


def search_optimum(kernel, experiments, folder, domain=None):
    from .optimizer import optimize_program
    program = model_program(kernel, kernel, domain=domain)
    optimize_program(folder + program["NAME"], program)
    experiments[program["NAME"]] = program


def explore_space(kernel, experiments, folder, domain=None, groupings=None, samples=SAMPLES, seed=None, suffixes=None):
    """OPT / HAND / AUTO / MAX / MIN, then ``samples`` random groupings (plus the five special ones), each with its
    six tile-count neighbours, deduplicated by TILING (ref diffusion.py:137-283).

    ``groupings`` replaces the random sample by given group-index lists (``grouping_of`` extracts them from the
    variant names the reference ships, whose own sample was unseeded); ``seed`` makes the sample reproducible;
    ``suffixes`` restricts the neighbours (default: all of mx my mz px py pz)."""
    import random
    sequence = P.KERNELS[kernel][1]()
    special = {
        "OPT": model_program(kernel, "OPT", domain=domain),
        "HAND": special_program(kernel, "HAND", domain),
        "AUTO": special_program(kernel, "AUTO", domain),
        "MAX": model_program(kernel, "MAX", groups=[(s, 0) for s in sequence], domain=domain),
        "MIN": model_program(kernel, "MIN", groups=[(s, n) for n, s in enumerate(sequence)], domain=domain),
    }
    for key, program in solve_all([(key, folder, program) for key, program in special.items()]):
        if "TILING" in program:
            experiments[key] = program
    # the optimum's own grouping becomes a constraint of the exploration (ref :151-155)
    if "OPT" in experiments:
        experiments["OPT"]["CONSTRAINTS"]["GROUPS"] = [
            (s, index) for index, outer in enumerate(experiments["OPT"]["TILING"]["GROUPS"])
            for s in outer["GROUPS"][0]["STENCILS"]]
    if groupings is None:
        rng = random.Random(seed if seed is not None else os.environ.get("ABSINTHE_SEED"))
        candidates = enumerate_groupings(sequence)
        groupings = rng.sample(candidates, min(samples, len(candidates)))
    groupings = [list(g) for g in groupings]
    for program in experiments.values():
        digits = [index for _, index in program["CONSTRAINTS"]["GROUPS"]]
        if digits not in groupings:
            groupings.append(digits)
    base = {}
    for digits in groupings:
        key = "%s-%s" % (kernel, "-".join(str(d) for d in digits))
        base[key] = model_program(kernel, key, groups=list(zip(sequence, digits)), name=key, domain=domain)
    exploration = {}
    for key, program in solve_all([(key, folder, program) for key, program in base.items()]):
        if "TILING" in program:
            exploration[key] = program
    jobs = []
    for key, solved in exploration.items():
        for suffix in suffixes if suffixes is not None else ("mx", "my", "mz", "px", "py", "pz"):
            name = "%s-%s" % (key, suffix)
            jobs.append((name, folder, model_program(
                kernel, name, groups=solved["CONSTRAINTS"]["GROUPS"], name=name, domain=domain,
                tiling=neighbour_tiling(solved, solved, suffix))))
    for key, program in solve_all(jobs):
        if "TILING" in program:
            exploration[key] = program
    for key, program in exploration.items():          # remove duplicates (ref :276-283)
        if any(existing["TILING"] == program["TILING"] for existing in experiments.values()):
            print("-> found existing experiment " + key)
        else:
            experiments[key] = program
    store_objectives(experiments, folder)


def store_objectives(experiments, folder):
    """``estimates.csv``: VAR, EST (ref diffusion.py:286-295)."""
    with open(folder + "estimates.csv", "w") as handle:
        out = csv.writer(handle, delimiter=",", quotechar="'", lineterminator="\n")
        out.writerow(["VAR", "EST"])
        for program in experiments.values():
            out.writerow([program["VARIANT"], str(program["OBJECTIVE"])])


def training_experiments(family, cores=P.B200_SMS, stride=20):
    """All (variant, tile shape) training programs (ref fitcache.py:116-172, fitddr.py:187-327)."""
    variants = S.FITCACHE_VARIANTS if family == "fitcache" else S.FITDDR_VARIANTS
    experiments = {}
    for variant in variants:
        for shape in P.training_tile_shapes():
            program = P.training_program(family, variant, shape, cores, stride)
            domain = "%dx%dx%d" % (program["X"], program["Y"], program["Z"])
            experiments["%s-%s" % (variant, domain)] = program
    return experiments


def main(kernel, arguments):
    """Entry point of the root-level scripts."""
    training = kernel in ("fitcache", "fitddr")
    optimize = explore = auto = generate = build = False
    parse, folder, stride = None, "./", 20
    try:
        # -s/--stride is an addition: execute every N-th training tile (reference protocol: 20)
        short = "gbvp:f:s:" if training else "oeagbp:f:"
        extended = (["generate", "build", "verify", "parse=", "folder=", "stride="] if training else
                    ["optimize", "explore", "auto", "generate", "build", "parse=", "folder="])
        opts, _ = getopt(arguments, short, extended)
    except GetoptError:
        print(kernel + ".py -f <folder>")
        sys.exit(2)
    for opt, arg in opts:
        if opt in ("-o", "--optimize"):
            optimize = True
        elif opt in ("-e", "--explore"):
            explore = True
        elif opt in ("-a", "--auto"):
            auto = True
        elif opt in ("-g", "--generate"):
            generate = True
        elif opt in ("-b", "--build"):
            build = True
        elif opt in ("-p", "--parse"):
            parse = arg
        elif opt in ("-f", "--folder"):
            folder = os.path.normpath(arg) + "/"
        elif opt in ("-s", "--stride"):
            stride = int(arg)
    print("-> working dir: " + str(folder))
    print("-> run generation: " + str(generate))
    print("-> run compilation: " + str(build))
    os.makedirs(folder, exist_ok=True)
    experiments = {}
    if training:
        experiments = training_experiments(kernel, stride=stride)
        template, cores = "template_training.cu", P.B200_SMS
    else:
        template, cores = "template_tiling.cu", P.base_program(kernel)["MACHINE"]["CORES"]
        if explore:
            explore_space(kernel, experiments, folder)
        elif optimize:
            search_optimum(kernel, experiments, folder)
        elif auto:
            auto_tune(kernel, experiments)
    if generate:
        for name, program in experiments.items():
            generate_code(template, folder + name + ".cu", deepcopy(program))
        generate_makefile(folder + "Makefile", experiments)
        generate_script(folder + "run.sh", experiments, cores, 1)
    if build:
        cwd = os.getcwd()
        os.chdir(folder)
        try:
            build_experiment(folder)
        finally:
            os.chdir(cwd)
    if parse is not None:
        if auto:
            write_results(parse_results(folder + parse, 0), folder + "auto.csv")
        else:
            write_results(parse_results(folder + parse, 16), folder + "results.csv")
    return experiments
