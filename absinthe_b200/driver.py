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

``-o`` and ``-e`` evaluate the reference's MILP model; they import ``stencil_optimizer`` from a
reference checkout (``ABSINTHE_REFERENCE`` or ``/root/reference``) and solve it with the ``cplex``
shim of ``tools/cplex`` -- that component is control plane and deliberately not rebuilt here.
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


def _reference_optimizer():
    root = os.environ.get("ABSINTHE_REFERENCE", "/root/reference")
    if not os.path.isfile(os.path.join(root, "stencil_optimizer.py")):
        raise SystemExit("-o/-e need the reference's stencil_optimizer.py (set ABSINTHE_REFERENCE) and the "
                         "cplex shim of tools/ on PATH")
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    os.environ["PATH"] = shim + os.pathsep + os.environ.get("PATH", "")
    sys.path.insert(0, root)
    try:
        import stencil_optimizer
    finally:
        sys.path.remove(root)
    return stencil_optimizer.optimize_program


def search_optimum(kernel, experiments, folder):
    program = P.base_program(kernel)
    program["STENCILS"] = P.KERNELS[kernel][0]()
    program["SEQUENCE"] = P.KERNELS[kernel][1]()
    _reference_optimizer()(folder + program["NAME"], program)
    experiments[program["NAME"]] = program


def explore_space(kernel, experiments, folder):
    """OPT, MAX and MIN of the model (the reference's exploration needs the same solver)."""
    optimize = _reference_optimizer()
    sequence = P.KERNELS[kernel][1]()
    constraints = {"OPT": None, "MAX": [(s, 0) for s in sequence],
                   "MIN": [(s, n) for n, s in enumerate(sequence)]}
    for name, groups in constraints.items():
        program = P.base_program(kernel)
        program["STENCILS"] = P.KERNELS[kernel][0]()
        program["SEQUENCE"] = sequence
        program["VARIANT"] = name
        if groups is not None:
            program["CONSTRAINTS"]["GROUPS"] = groups
        optimize(folder + name, program)
        experiments[name] = program
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
