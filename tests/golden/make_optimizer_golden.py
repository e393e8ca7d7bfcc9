"""Generate tests/golden/optimizer.json from the REAL reference (development container only).

For a spread of programs (the three sequences; OPT / MAX / MIN / tile-count constraints; the reference's
CPU parameters and the B200 ones) the reference's own ``stencil_optimizer.generate_lp`` writes its LP;
the LP is parsed (``absinthe_b200.milp.read_lp``) into a canonical, order-independent form and its sha256
recorded, together with row / variable counts.  Cases marked ``solve`` are also solved with HiGHS and
their objective recorded.  ``tests/test_optimizer.py`` rebuilds the same programs with
``absinthe_b200.optimizer`` and must arrive at the same fingerprints and objectives.

    python tests/golden/make_optimizer_golden.py
"""

import hashlib
import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stdout
from copy import deepcopy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden", "optimizer.json")


def cases():
    """name -> (kernel, parameters 'cpu' | 'b200', constraints, solve?)"""
    from absinthe_b200 import programs as P
    out = {}
    for kernel in ("diffusion", "advection", "fastwaves"):
        seq = P.KERNELS[kernel][1]()
        fused = [(s, 0) for s in seq]
        split = [(s, n) for n, s in enumerate(seq)]
        out[kernel + "-opt-cpu"] = (kernel, "cpu", {}, False)
        out[kernel + "-max-cpu"] = (kernel, "cpu", {"GROUPS": fused}, kernel == "fastwaves")
        out[kernel + "-min-cpu"] = (kernel, "cpu", {"GROUPS": split}, kernel == "advection")
        out[kernel + "-max-b200"] = (kernel, "b200", {"GROUPS": fused}, kernel != "diffusion")
        out[kernel + "-tiling-cpu"] = (kernel, "cpu", {"GROUPS": fused, "TILING": [("x", seq[0], -3), ("z", seq[-1], 4)]},
                                       False)
    return out


def program_of(kernel, params, constraints):
    from absinthe_b200 import programs as P
    if params == "cpu":
        program = P.base_program(kernel, machine={"CORES": 4, "CAPACITY": 85 * 1024}, memory=P.CPU_MEMORY,
                                 cache=P.CPU_CACHE)
    else:
        # B200 machine (148 SMs, 227 KiB) with the round-1 GPU fit, pinned here so that later re-fits of
        # programs.B200_MEMORY / B200_CACHE do not move the fixtures
        program = P.base_program(kernel, memory={"RW BODY": 0.0, "ST BODY": 4.91e-9, "RW PEEL": 0.0, "ST PEEL": 2.75e-8},
                                 cache={"BODY": 1.17e-10, "PEEL": 2.33e-10})
    program["STENCILS"] = P.KERNELS[kernel][0]()
    program["SEQUENCE"] = P.KERNELS[kernel][1]()
    program["CONSTRAINTS"] = deepcopy(constraints)
    return program


def digest(canonical):
    return hashlib.sha256(json.dumps(canonical, sort_keys=True).encode()).hexdigest()


def main():
    from absinthe_b200 import milp
    sys.path.insert(0, REFERENCE)
    import stencil_optimizer as R
    golden = {}
    for name, (kernel, params, constraints, solve) in sorted(cases().items()):
        program = program_of(kernel, params, constraints)
        with tempfile.TemporaryDirectory() as tmp, redirect_stdout(io.StringIO()):
            R.compute_dependencies(program)
            R.compute_sequence(program)
            R.compute_utilization(program)
            R.compute_domain(program)
            R.generate_lp(os.path.join(tmp, "m"), program)
            model = milp.read_lp(os.path.join(tmp, "m.lp"))
        canonical = model.canonical()
        entry = {"sha256": digest(canonical), "rows": len(canonical["rows"]), "general": len(canonical["general"]),
                 "binary": len(canonical["binary"])}
        if solve:
            entry["objective"] = model.solve(time_limit=300).objective
        golden[name] = entry
        print(name, entry, flush=True)
    with open(GOLDEN, "w") as handle:
        json.dump(golden, handle, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
