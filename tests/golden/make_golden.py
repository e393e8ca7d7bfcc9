"""Generate tests/golden/* from the REAL reference (development container only).

Run as ``PYTHONHASHSEED=0 python tests/golden/make_golden.py``.  Needs ``/root/reference``; the
fixtures it writes are small and committed, so the tests that use them run anywhere:

* ``stencil_fingerprints.json``  whitespace-insensitive sha256 prefix of every stencil string of
                                 the reference drivers (diffusion/advection/fastwaves/fitcache/fitddr)
* ``analysis.json``              INPUTS/OUTPUTS/TEMPS/LOOPS/HALOS/ID produced by the reference's
                                 ``stencil_analyzer`` for a spread of groupings and tilings
* ``reference_digests.json``     sha256 of the inputs the reference's ``main()`` would generate
                                 (srand(0)/rand(), periodic) and of the outputs its own
                                 ``compute_reference`` computes from them, strict IEEE build
* ``parse_results.json``         rows the reference's ``parse_results`` extracts from a sample log
* ``training_digests.json``     the same for four fitcache / fitddr training programs
* ``shipped_medians.json``       estimate and median measurement of every variant the reference ships
                                 results for (``*/estimates.csv`` joined with ``*/results.csv``)
"""

import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def fingerprint(code):
    return hashlib.sha256("".join(code.split()).encode()).hexdigest()[:16]


def stencil_fingerprints():
    sys.path.insert(0, REFERENCE)
    import advection
    import diffusion
    import fastwaves
    import fitcache
    import fitddr
    out = {}
    for module in (diffusion, advection, fastwaves):
        out[module.PROGRAM["NAME"]] = {k: fingerprint(v) for k, v in module.STENCILS.items()}
    for name in ("PT8", "PT12", "PT16", "PT20"):
        out[name] = {k: fingerprint(v) for k, v in getattr(fitcache, name).items()}
    for name in ("IN0BD0", "IN1BD0", "IN2BD0", "IN3BD0", "IN1BD1", "IN2BD1", "IN3BD1", "IN1BD2", "IN2BD2",
                 "IN3BD2"):
        out[name] = {k: fingerprint(v) for k, v in getattr(fitddr, name).items()}
    sys.path.remove(REFERENCE)
    return out


def normalise(tiling):
    def group(g):
        d = {k: sorted(g[k]) for k in ("INPUTS", "OUTPUTS", "TEMPS")}
        d["LOOPS"] = {k: [list(x) for x in v] for k, v in sorted(g["LOOPS"].items())}
        d["HALOS"] = {k: {a: list(b) for a, b in sorted(v.items())} for k, v in sorted(g["HALOS"].items())}
        d["ID"] = g["ID"]
        if "GROUPS" in g:
            d["GROUPS"] = [group(x) for x in g["GROUPS"]]
        return d
    return {"INPUTS": sorted(tiling["INPUTS"]), "OUTPUTS": sorted(tiling["OUTPUTS"]),
            "TEMPS": sorted(tiling["TEMPS"]), "GROUPS": [group(g) for g in tiling["GROUPS"]]}


def analysis_cases():
    from absinthe_b200 import programs as P
    rng = random.Random(2019)
    cases = []
    for kernel in ("diffusion", "advection", "fastwaves"):
        seq = P.KERNELS[kernel][1]()
        assignments = [[0] * len(seq), list(range(len(seq)))]
        for _ in range(10):
            a = [0]
            for _ in seq[1:]:
                a.append(a[-1] + rng.choice([0, 0, 1]))
            assignments.append(a)
        for a in assignments:
            groups = P.split_sequence(seq, a)
            tiles = [[rng.choice([1, 2, 4, 8]), rng.choice([1, 2, 4, 16]), rng.choice([1, 5, 15, 60])]
                     for _ in groups]
            cases.append({"kernel": kernel, "assignment": a, "tiles": tiles})
    return cases


def analysis_fixtures():
    sys.path.insert(0, REFERENCE)
    import stencil_analyzer as R
    from absinthe_b200 import programs as P
    out = []
    for case in analysis_cases():
        seq = P.KERNELS[case["kernel"]][1]()
        program = P.make_program(case["kernel"], P.split_sequence(seq, case["assignment"]),
                                 [tuple(t) for t in case["tiles"]])
        R.verify_program(program)
        R.compute_dataflow(program)
        R.compute_boundaries(program)
        case = dict(case)
        case["expected"] = normalise(program["TILING"])
        out.append(case)
    sys.path.remove(REFERENCE)
    return out


def reference_digests():
    import numpy as np
    from common import zoo
    from oracle import build_ref

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    out = {}
    for name, program in zoo():
        dump = build_ref.build_dump(program, "golden_" + name)
        inputs = dump.fill_inputs()
        outputs = dump.compute(inputs)
        out[name] = {
            "input_order": dump.inputs,
            "inputs": {k: sha(v) for k, v in inputs.items()},
            "outputs": {k: sha(v) for k, v in outputs.items()},
            "abs_max": {k: float(np.abs(v).max()) for k, v in outputs.items()},
        }
        print("digest", name)
    return out


TRAINING_CASES = [("fitcache", "PT12", (10, 3, 4)), ("fitcache", "PT20", (10, 2, 5)), ("fitddr", "IN2BD1", (10, 3, 4)),
                  ("fitddr", "IN3BD2", (20, 2, 3))]


def training_digests():
    """Reference-computed values for training programs: the program of fitcache.py / fitddr.py (all nine stencils of a
    variant, 2 x 2 x 5*CORES tiles) rendered by the reference's generator with template_tiling.cpp, whose
    compute_reference evaluates the same stencils sequentially over the whole domain, strict flags."""
    import numpy as np
    from absinthe_b200 import programs as P
    from oracle import build_ref

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    out = {}
    for family, variant, shape in TRAINING_CASES:
        program = P.training_program(family, variant, shape, cores=2, stride=20)
        program.pop("TRAINING")
        name = "%s-%s-%s" % (family, variant, "x".join(map(str, shape)))
        dump = build_ref.build_dump(program, "golden_" + name.replace("-", "_"))
        inputs = dump.fill_inputs()
        outputs = dump.compute(inputs)
        out[name] = {"input_order": dump.inputs, "inputs": {k: sha(v) for k, v in inputs.items()},
                     "outputs": {k: sha(v) for k, v in outputs.items()},
                     "abs_max": {k: float(np.abs(v).max()) for k, v in outputs.items()}}
        print("digest", name)
    return out


SAMPLE_LOG = """-> configuration
   - variant alpha
   - domain 64, 64, 60
   - runs 3
-> apply stencils...
   - total time (min/median/max) [ms]: 2.5
   - halo time (min/median/max) [ms]: 0.5
-> apply stencils...
   - total time (min/median/max) [ms]: 2.25
   - halo time (min/median/max) [ms]: 0.125
-> apply stencils...
   - total time (min/median/max) [ms]: 1.75
   - halo time (min/median/max) [ms]: 0.25
-> configuration
   - variant beta
   - domain 32, 16, 8
   - runs 2
   - total time (min/median/max) [ms]: 0.5
   - halo time (min/median/max) [ms]: 0
   - total time (min/median/max) [ms]: 0.375
   - halo time (min/median/max) [ms]: 0
"""


def parse_fixture():
    sys.path.insert(0, REFERENCE)
    import stencil_generator as R
    path = os.path.join(GOLDEN, "sample_output.txt")
    with open(path, "w") as handle:
        handle.write(SAMPLE_LOG)
    out = {"skip0": R.parse_results(path, 0), "skip1": R.parse_results(path, 1), "skip16": R.parse_results(path)}
    sys.path.remove(REFERENCE)
    return out


def shipped_variants():
    """VAR names of the implementation variants the reference ships estimates for (*/estimates.csv)."""
    import csv
    out = {}
    for kernel in ("diffusion", "advection", "fastwaves"):
        with open(os.path.join(REFERENCE, kernel, "estimates.csv")) as handle:
            out[kernel] = sorted(r["VAR"] for r in csv.DictReader(handle))
    return out


def shipped_medians():
    """VAR -> [EST, median(TOTAL - HALO)] of the reference's shipped estimates.csv / results.csv."""
    import csv
    import statistics
    out = {}
    for kernel in ("diffusion", "advection", "fastwaves"):
        with open(os.path.join(REFERENCE, kernel, "estimates.csv")) as handle:
            est = {r["VAR"]: float(r["EST"]) for r in csv.DictReader(handle)}
        mea = {}
        with open(os.path.join(REFERENCE, kernel, "results.csv")) as handle:
            for r in csv.DictReader(handle):
                mea.setdefault(r["VAR"], []).append(float(r["TOTAL"]) - float(r["HALO"]))
        out[kernel] = {var: [est[var], statistics.median(v)] for var, v in sorted(mea.items()) if var in est}
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name, fn in (("stencil_fingerprints", stencil_fingerprints), ("analysis", analysis_fixtures),
                     ("parse_results", parse_fixture), ("shipped_variants", shipped_variants),
                     ("shipped_medians", shipped_medians),
                     ("reference_digests", reference_digests), ("training_digests", training_digests)):
        with open(os.path.join(GOLDEN, name + ".json"), "w") as handle:
            json.dump(fn(), handle, indent=1, sort_keys=True)
        print("wrote", name)


if __name__ == "__main__":
    main()
