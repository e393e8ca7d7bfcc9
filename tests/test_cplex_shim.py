"""The cplex stand-in (tools/cplex): LP-format parsing, MILP solve, XML solution the reference parses."""

import os
import subprocess
import sys
from xml.dom.minidom import parse

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tools", "cplex")

LP = """Minimize
3 x + 2 y + t%0 - -0.5 z
Subject To
\\ a comment line
x + y >= 3.5
x - y <= 1
t%0 - 2 x >= 0
n#a + n#b = 1
z - 4 n#a - 7 n#b >= 0
General
x y
Binary
n#a n#b
End
"""


def run_shim(tmp_path, text):
    lp = tmp_path / "m.lp"
    lp.write_text(text)
    sol = tmp_path / "m.sol"
    script = "read %s\nmipopt\nwrite %s\nquit\n" % (lp, sol)
    subprocess.run([sys.executable, SHIM], input=script.encode(), check=True, capture_output=True)
    dom = parse(str(sol))
    values = {v.attributes["name"].value: float(v.attributes["value"].value)
              for v in dom.getElementsByTagName("variable")}
    return float(dom.getElementsByTagName("header")[0].attributes["objectiveValue"].value), values


def test_small_milp(tmp_path):
    objective, v = run_shim(tmp_path, LP)
    # integers x + y >= 3.5 -> x + y = 4 with x - y <= 1: (x, y) = (2, 2) costs 6 + 4 + t(=4) = 14;
    # (1, 3) costs 3 + 6 + 2 = 11; (0, 4) costs 8.  "- -0.5 z" is + 0.5 z with z >= 4 -> + 2.
    assert (v["x"], v["y"]) == (0.0, 4.0)
    assert v["n#a"] == 1.0 and v["n#b"] == 0.0 and v["z"] == pytest.approx(4.0)
    assert objective == pytest.approx(8.0 + 2.0)


@pytest.mark.skipif(not os.path.isfile("/root/reference/stencil_optimizer.py"), reason="reference not mounted")
def test_reproduces_a_shipped_estimate(tmp_path):
    """HAND-tuned fast waves: the reference ships its estimate (fastwaves/estimates.csv, VAR = HAND)."""
    import csv
    from copy import deepcopy
    shipped = {r["VAR"]: float(r["EST"]) for r in csv.DictReader(open("/root/reference/fastwaves/estimates.csv"))}
    sys.path.insert(0, "/root/reference")
    env_path = os.environ.get("PATH", "")
    os.environ["PATH"] = os.path.join(ROOT, "tools") + os.pathsep + env_path
    os.environ["ABSINTHE_MILP_SECONDS"] = "60"
    try:
        import fastwaves as F
        import stencil_optimizer as O
        seq = ["ppgk", "ppgc", "ppgu", "ppgv", "uout", "vout", "udc", "vdc", "div"]
        program = deepcopy(F.PROGRAM)
        program.update(STENCILS=F.STENCILS, VARIANT="MAX", SEQUENCE=seq)
        program["CONSTRAINTS"]["GROUPS"] = [(s, 0) for s in seq]
        O.optimize_program(str(tmp_path / "MAX"), program)
        assert float(program["OBJECTIVE"]) == pytest.approx(shipped["MAX"], rel=2e-4)
    finally:
        sys.path.remove("/root/reference")
        os.environ["PATH"] = env_path
        for name in ("fastwaves", "stencil_optimizer", "stencil_generator", "stencil_analyzer"):
            sys.modules.pop(name, None)
