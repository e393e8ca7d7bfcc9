"""The experiment drivers keep the reference's command-line recipe (-a/-g/-b/-p/-f)."""

import os
import subprocess
import sys

import pytest

from absinthe_b200 import driver
from absinthe_b200 import generator as G
from absinthe_b200 import programs as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_auto_generate_writes_variants_makefile_and_script(tmp_path):
    folder = str(tmp_path / "advection")
    experiments = driver.main("advection", ["-a", "-g", "-f", folder])
    # ref advection.py:102-107: nx, ny in 1..64, nz in {1,2,5,15,30,60}, at least CORES (= 148 SMs) tiles
    assert len(experiments) == 198 and all(n.startswith("advection-0-") for n in experiments)
    for name in list(experiments)[:3]:
        assert os.path.exists(os.path.join(folder, name + ".cu")) and os.path.exists(os.path.join(folder, name + ".desc"))
    make = open(os.path.join(folder, "Makefile")).read()
    assert "advection-0-64-64-60.cubin: advection-0-64-64-60.cu" in make
    # the Makefile rule really builds a cubin for sm_100a
    target = "advection-0-4-8-5.cubin"
    subprocess.check_call(["make", target], cwd=folder, stdout=subprocess.DEVNULL)
    assert os.path.getsize(os.path.join(folder, target)) > 1000
    script = open(os.path.join(folder, "run.sh")).read()
    assert "absinthe_b200.runner" in script and "./advection-0-4-8-5" in script and "PYTHONPATH=" + ROOT in script


def test_fastwaves_auto_is_unfiltered_like_the_reference():
    experiments = {}
    driver.auto_tune("fastwaves", experiments)
    assert len(experiments) == 2 * 7 * 7 * 6          # ref fastwaves.py:100: the >= CORES filter is commented out
    assert experiments["fastwaves-1-1-1-1"]["OUTPUTS"] == ["div"]


def test_parse_flag_writes_auto_csv(tmp_path):
    folder = tmp_path / "diffusion"
    folder.mkdir()
    log = open(os.path.join(ROOT, "tests", "golden", "sample_output.txt")).read()
    (folder / "output.txt").write_text(log)
    driver.main("diffusion", ["-a", "-p", "output.txt", "-f", str(folder)])
    rows = (folder / "auto.csv").read_text().splitlines()
    assert rows[0] == "VAR,X,Y,Z,TOTAL,HALO" and len(rows) == 6
    driver.main("diffusion", ["-p", "output.txt", "-f", str(folder)])          # default skip = 16 -> no rows
    assert (folder / "results.csv").read_text().splitlines() == ["VAR,X,Y,Z,TOTAL,HALO"]


def test_training_drivers_enumerate_the_reference_sweep():
    cache = driver.training_experiments("fitcache", cores=4)
    ddr = driver.training_experiments("fitddr", cores=4)
    assert len(cache) == 4 * 103 and len(ddr) == 10 * 103       # ref fitcache.py:116-172, fitddr.py:187-327
    p = cache["PT12-100x10x60"]          # the first sample of the reference's fitcache/results.csv
    assert (p["X"], p["Y"], p["Z"]) == (100, 10, 60) and p["TILING"]["GROUPS"][0]["GROUPS"][0]["NZ"] == 20


def test_training_stride_option(tmp_path):
    """-s/--stride: every N-th tile is executed (20 = the reference's one-tile-per-core protocol)."""
    from absinthe_b200 import generator as G
    default = driver.training_experiments("fitcache", cores=4)["PT12-100x10x60"]
    assert default["TRAINING"]["STRIDE"] == 20
    full = driver.training_experiments("fitcache", cores=4, stride=1)["PT12-100x10x60"]
    assert full["TRAINING"]["STRIDE"] == 1
    G.generate_code("template_training.cu", str(tmp_path / "t.cu"), full)
    assert " training_step 1 " in (tmp_path / "t.desc").read_text()


def test_root_scripts_run(tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "fitcache.py"), "-f", str(tmp_path)],
                         capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "-> working dir" in out.stdout


# ---- the -e exploration space (ref diffusion.py:137-283) ------------------------------------------
def test_grouping_enumeration_matches_the_reference_counts():
    from absinthe_b200 import driver as D
    # 2^(n-1) assignments of consecutive stencils, those with a group index above 4 dropped (ref :207-221)
    assert len(D.enumerate_groupings(P.KERNELS["fastwaves"][1]())) == 1 + 8 + 28 + 56 + 70
    assert len(D.enumerate_groupings(P.KERNELS["diffusion"][1]())) == 1 + 15 + 105 + 455 + 1365
    assert [0, 0, 0, 0, 0, 0, 1, 1, 1] in D.enumerate_groupings(P.KERNELS["fastwaves"][1]())


def test_hand_and_auto_constraints_are_the_reference_lists():
    from absinthe_b200 import driver as D
    hand = D.special_program("diffusion", "HAND")
    assert set(hand["CONSTRAINTS"]["TILING"]) == {(d, s, v) for s in hand["SEQUENCE"]
                                                   for d, v in (("x", -2), ("y", -2), ("z", 59))}
    assert hand["MACHINE"]["CAPACITY"] == 2 ** 31 and hand["SLACK"] == {"SIZE": 1.0, "CORES": 1.0}
    auto = D.special_program("advection", "AUTO")
    assert set(auto["CONSTRAINTS"]["TILING"]) == {(d, s, v) for s in auto["SEQUENCE"]
                                                   for d, v in (("x", -2), ("y", -2), ("z", 29), ("z", -31))}
    fw = D.special_program("fastwaves", "AUTO")
    assert ("y", "ppgk", -5) in fw["CONSTRAINTS"]["TILING"] and ("y", "div", 1) in fw["CONSTRAINTS"]["TILING"]
    assert dict(fw["CONSTRAINTS"]["GROUPS"])["udc"] == 1
    assert fw["MACHINE"]["CORES"] == 10          # 1 x 2 x 5 tiles in the second group: fewer than SMs


def test_shipped_variant_names_decode():
    import json
    import os
    from absinthe_b200 import driver as D
    names = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "shipped_variants.json")))
    for kernel, variants in names.items():
        every = D.enumerate_groupings(P.KERNELS[kernel][1]())
        for name in variants:
            if name.startswith(kernel + "-"):
                digits, suffix = D.grouping_of(kernel, name)
                # sampled groupings have at most five groups; the special ones (MIN: one group per stencil) are
                # explored whatever their group count (ref diffusion.py:225-226)
                assert digits in every or digits == list(range(len(digits))), name
                assert suffix in (None, "mx", "my", "mz", "px", "py", "pz"), name


def test_neighbour_tiling_bounds():
    from absinthe_b200 import driver as D
    solved = {"X": 64, "Y": 64, "Z": 60, "TILING": {"GROUPS": [
        {"GROUPS": [{"STENCILS": ["a", "b"], "NX": 1, "NY": 8, "NZ": 60}]},
        {"GROUPS": [{"STENCILS": ["c"], "NX": 4, "NY": 64, "NZ": 2}]}]}}
    assert D.neighbour_tiling(solved, solved, "mx") == [("x", "c", -4)]                       # n <= 3; NX = 1 stays
    assert D.neighbour_tiling(solved, solved, "py") == [("y", "a", 8), ("y", "b", 8)]         # n >= 9; NY = 64 stays
    assert D.neighbour_tiling(solved, solved, "pz") == [("z", "c", 2)]


def test_small_exploration_end_to_end(tmp_path, monkeypatch):
    """-e on advection with one sampled grouping and one neighbour kind: estimates.csv, dedup, generation."""
    from absinthe_b200 import driver as D
    monkeypatch.setenv("ABSINTHE_MILP_SECONDS", "10")
    experiments = {}
    D.explore_space("advection", experiments, str(tmp_path) + "/", samples=1, seed=3, suffixes=("pz",))
    assert {"OPT", "HAND", "AUTO", "MAX"} <= set(experiments)
    assert [g["GROUPS"][0]["NZ"] for g in experiments["AUTO"]["TILING"]["GROUPS"]] == [30]
    assert [tuple(g["GROUPS"][0][k] for k in ("NX", "NY", "NZ")) for g in experiments["HAND"]["TILING"]["GROUPS"]] \
        == [(1, 1, 60)]
    rows = (tmp_path / "estimates.csv").read_text().splitlines()
    assert rows[0] == "VAR,EST" and len(rows) == 1 + len(experiments)
    tilings = [str(p["TILING"]) for p in experiments.values()]
    assert len(set(tilings)) == len(tilings)                                                 # duplicates removed
    assert all(k in ("OPT", "HAND", "AUTO", "MAX", "MIN") or k.startswith("advection-") for k in experiments)
    for program in experiments.values():                                                     # every variant renders
        source, _, _ = G.render("template_tiling.cu", P.clone(program))
        assert "absinthe_advection" in source
