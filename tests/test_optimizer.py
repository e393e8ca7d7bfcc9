"""The analytical model as a MILP (absinthe_b200/optimizer.py): every row equals the reference's.

``tests/golden/optimizer.json`` holds, per case, the fingerprint of the canonical form of the LP the
reference's own ``stencil_optimizer.generate_lp`` writes (made by ``tests/golden/make_optimizer_golden.py``
in the development container) and, for the cases solved there, the optimum.  No reference checkout is
needed to run these tests."""

import io
import json
import os
from contextlib import redirect_stdout

import pytest

from absinthe_b200 import milp
from absinthe_b200 import optimizer as O
from golden.make_optimizer_golden import cases, digest, program_of

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "optimizer.json")))
CASES = cases()


def analysed(name):
    kernel, params, constraints, _ = CASES[name]
    program = program_of(kernel, params, constraints)
    with redirect_stdout(io.StringIO()):
        O.compute_dependencies(program)
        O.compute_sequence(program)
        O.compute_utilization(program)
        O.compute_domain(program)
    return program


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_model_equals_the_reference_lp(name, tmp_path):
    model = O.build_model(analysed(name))
    canonical = model.canonical()
    want = GOLDEN[name]
    assert (len(canonical["rows"]), len(canonical["general"]), len(canonical["binary"])) == (
        want["rows"], want["general"], want["binary"])
    assert digest(canonical) == want["sha256"]
    # ... and the LP text this module writes reads back as the same model
    model.write_lp(str(tmp_path / "m.lp"))
    assert digest(milp.read_lp(str(tmp_path / "m.lp")).canonical()) == want["sha256"]


@pytest.mark.parametrize("name", sorted(n for n in GOLDEN if "objective" in GOLDEN[n] and n != "advection-min-cpu"))
def test_optimum_matches(name, tmp_path):
    program = analysed(name)
    with redirect_stdout(io.StringIO()):
        model = O.generate_lp(str(tmp_path / "m"), program)
        O.solve_lp(str(tmp_path / "m"), model)
        O.parse_lp(str(tmp_path / "m"), program)
    assert float(program["OBJECTIVE"]) == pytest.approx(GOLDEN[name]["objective"], rel=2e-4)
    # the cost model re-evaluated on the integer part of the solution (the check the reference prints): the
    # objective can only be larger (helper variables need not be tight where they do not pay)
    assert 0.5 * float(program["OBJECTIVE"]) < program["ESTIMATE"]["total"] <= float(program["OBJECTIVE"]) * (1 + 1e-6)
    groups = program["TILING"]["GROUPS"]
    assert len(groups) == 1 and sorted(groups[0]["GROUPS"][0]) == ["NX", "NY", "NZ", "STENCILS"]


def test_shipped_estimate_of_fused_fast_waves():
    """The reference ships EST(MAX) of fast waves for its CPU parameters (fastwaves/estimates.csv)."""
    shipped = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "shipped_medians.json")))
    assert GOLDEN["fastwaves-max-cpu"]["objective"] == pytest.approx(shipped["fastwaves"]["MAX"][0], rel=2e-4)


def test_random_sequence_respects_dependencies():
    kernel, params, constraints, _ = CASES["fastwaves-opt-cpu"]
    program = program_of(kernel, params, constraints)
    del program["SEQUENCE"]
    with redirect_stdout(io.StringIO()):
        O.compute_dependencies(program)
        O.compute_sequence(program)
    order = program["SEQUENCE"]
    assert sorted(order) == sorted(program["STENCILS"])
    assert order.index("ppgk") < order.index("ppgc") < order.index("ppgu") < order.index("uout") < order.index("udc")


def test_small_model_round_trip(tmp_path):
    m = milp.Model("toy")
    m.minimize([(3, "x"), (2, "y"), (1, "t%0"), (0.5, "z")])
    m.general("x", "y")
    m.binary("n#a", "n#b")
    m.add([(1, "x"), (1, "y")], ">=", 3.5)
    m.add([(1, "x"), (-1, "y")], "<=", 1)
    m.add([(1, "t%0"), (-2, "x")], ">=", 0)
    m.add([(1, "n#a"), (1, "n#b")], "=", 1)
    m.add([(1, "z"), (-4, "n#a"), (-7, "n#b")], ">=", 0)
    s = m.solve()
    assert (s["x"], s["y"], s["n#a"]) == (0.0, 4.0, 1.0) and s.objective == pytest.approx(10.0)
    milp.write_solution(str(tmp_path / "toy.sol"), m, s)
    objective, values = milp.read_solution(str(tmp_path / "toy.sol"))
    assert float(objective) == pytest.approx(10.0) and values["y"] == 4.0
