"""The analyzer reproduces the reference's variant description (fixtures from the real
stencil_analyzer.py, see tests/golden/make_golden.py) plus the tile geometry rules."""

import json
import os

import pytest

from absinthe_b200 import analysis as A
from absinthe_b200 import programs as P
from absinthe_b200.emitter import c_trunc_div, tile_geometry

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "analysis.json")))


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


@pytest.mark.parametrize("case", CASES, ids=["%s-%d" % (c["kernel"], n) for n, c in enumerate(CASES)])
def test_variant_description_matches_reference(case):
    seq = P.KERNELS[case["kernel"]][1]()
    program = P.make_program(case["kernel"], P.split_sequence(seq, case["assignment"]),
                             [tuple(t) for t in case["tiles"]])
    A.analyse(program)
    assert json.loads(json.dumps(normalise(program["TILING"]))) == case["expected"]


def test_analyse_is_idempotent():
    program = P.make_program("fastwaves")
    A.analyse(program)
    snapshot = json.dumps(normalise(program["TILING"]))
    A.analyse(program)
    assert json.dumps(normalise(program["TILING"])) == snapshot
    assert len(program["TILING"]["GROUPS"]) == 2          # dummy + one fused group


def test_fetch_counts():
    # diffusion 6+5+5+7 per field
    d = P.KERNELS["diffusion"][0]()
    assert [A.count_fetches(d[n]) for n in ("ulap", "ufli", "uflj", "uout")] == [6, 5, 5, 7]
    a = P.KERNELS["advection"][0]()
    assert [A.count_fetches(a[n]) for n in P.KERNELS["advection"][1]()] == [3, 9, 5, 10, 5, 9, 3, 10]
    f = P.KERNELS["fastwaves"][0]()
    assert [A.count_fetches(f[n]) for n in P.KERNELS["fastwaves"][1]()] == [4, 3, 9, 9, 6, 6, 5, 5, 17]


def test_algorithmic_bytes():
    # SURVEY.md section 8d
    for kernel, fused, unfused in (("diffusion", 72, 416), ("advection", 32, 176), ("fastwaves", 104, 288)):
        seq = P.KERNELS[kernel][1]()
        assert A.algorithmic_bytes_per_point(A.analyse(P.make_program(kernel))) == fused
        single = P.make_program(kernel, [[s] for s in seq], [(1, 1, 1)] * len(seq))
        assert A.algorithmic_bytes_per_point(A.analyse(single)) == unfused
    f = P.KERNELS["fastwaves"][1]()
    assert A.algorithmic_bytes_per_point(A.analyse(P.make_program("fastwaves", [f[:6], f[6:]]))) == 128


def test_too_wide_access_is_rejected():
    program = P.make_program("advection")
    program["HX"] = 2
    with pytest.raises(AssertionError, match="accesses too wide"):
        A.analyse(program)


def test_tile_geometry_truncates_like_cpp():
    # ref template_tiling.cpp:228-233: T = ceil(S/N), O = -(T*N - S)/2 with C++ truncation
    assert tile_geometry(64, 8) == (8, 0)
    assert tile_geometry(60, 8) == (8, -2)
    assert tile_geometry(64, 60) == (2, -28)
    assert tile_geometry(37, 5) == (8, -1)       # -(40-37)/2 = -1 in C++, -2 with Python floor
    assert c_trunc_div(-3, 2) == -1 and c_trunc_div(3, 2) == 1 and c_trunc_div(-4, 2) == -2


def test_empty_tiles_rejected():
    program = P.make_program("advection")
    program["TILING"]["NX"] = 0
    with pytest.raises(Exception):
        A.verify_program(program)
