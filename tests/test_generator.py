"""Host logic of the generator surface: rendering, descriptor, Makefile/run.sh, results parsing."""

import json
import os

import pytest

from absinthe_b200 import generator as G
from absinthe_b200 import programs as P
from absinthe_b200.emitter import Plan, SMEM_LIMIT
from common import zoo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_parse_results_matches_reference_rows():
    want = json.load(open(os.path.join(GOLDEN, "parse_results.json")))
    sample = os.path.join(GOLDEN, "sample_output.txt")
    assert G.parse_results(sample, 0) == want["skip0"]
    assert G.parse_results(sample, 1) == want["skip1"]
    assert G.parse_results(sample) == want["skip16"]


def test_write_results_format(tmp_path):
    rows = G.parse_results(os.path.join(GOLDEN, "sample_output.txt"), 0)
    table = tmp_path / "results.csv"
    G.write_results(rows, str(table))
    lines = table.read_text().split("\n")
    assert lines[0] == "VAR,X,Y,Z,TOTAL,HALO"
    assert lines[1] == "alpha,64,64,60,2.5,0.5"


@pytest.mark.skipif(not os.path.isdir("/root/reference/diffusion"), reason="reference data not mounted")
def test_parse_results_on_shipped_log():
    import csv
    rows = G.parse_results("/root/reference/diffusion/output.txt")
    with open("/root/reference/diffusion/results.csv") as handle:
        shipped = list(csv.reader(handle, quotechar="'"))[1:]
    assert rows == shipped


def test_generate_code_writes_source_and_descriptor(tmp_path):
    program = P.make_program("fastwaves", variant="fw")
    G.generate_code("template_tiling.cu", str(tmp_path / "fw.cu"), program)
    source = (tmp_path / "fw.cu").read_text()
    desc = (tmp_path / "fw.desc").read_text()
    assert "absinthe_fastwaves_g1" in source and "cp.async.bulk.tensor.3d" in source
    assert desc.startswith("absinthe-variant 1\nname fw\n")
    assert "SCHEDULE" in program and [e["TYPE"] for e in program["SCHEDULE"]] == ["PUT", "WAIT", "COMP"]
    with pytest.raises(ValueError):
        G.generate_code("template_tiling.cpp", str(tmp_path / "x.cpp"), P.make_program("advection"))


def test_stencil_text_is_pasted_with_only_accesses_rewritten():
    import re
    program = P.make_program("advection")
    source, _, flags = G.render("template_tiling.cu", program)
    assert "-fmad=false" in flags
    strip = lambda s: re.sub(r"\s+", "", s)
    for name, code in program["STENCILS"].items():
        # rewriting accesses back to placeholders must give the original skeleton
        skeleton = strip(re.sub(r"\w+\([^()]*\)", "@", code))
        assert skeleton in strip(re.sub(r"\b[lv]_\w+", "@", source)), name


@pytest.mark.parametrize("name,program", zoo(), ids=[c[0] for c in zoo()])
def test_plans_fit_shared_memory_and_cover_tiles(name, program):
    plan = Plan(P.clone(program))
    for g in plan.groups:
        assert g.smem_bytes <= SMEM_LIMIT
        for a in range(3):
            assert g.B[a] <= g.T[a] and g.nsub[a] * g.B[a] >= g.T[a]
            assert g.T[a] * g.N[a] >= g.S[a]                       # tiles cover the domain
            assert g.O[a] <= 0 and g.T[a] * g.N[a] + g.O[a] >= g.S[a]
        for p in g.input_plan:
            assert max(p["DIM"]) <= 256                            # TMA box limit
            if g.load == "tma":                                    # 16-byte inner extent, room for the
                need = g.B[0] + g.reach[p["NAME"]][0][1] - g.reach[p["NAME"]][0][0]
                assert p["DIM"][0] % 2 == 0 and p["DIM"][0] >= need + 1   # even-column start shift


def test_makefile_and_script(tmp_path):
    experiments = {"a": P.make_program("advection"), "b": P.make_program("diffusion")}
    G.generate_makefile(str(tmp_path / "Makefile"), experiments)
    G.generate_script(str(tmp_path / "run.sh"), experiments, 148, 2)
    make = (tmp_path / "Makefile").read_text()
    assert "a.cubin: a.cu" in make and "arch=compute_100a,code=sm_100a" in make and "-fmad=false" in make
    script = (tmp_path / "run.sh").read_text()
    assert script.count("absinthe_b200.runner") == 2 and script.count("./a") == 2 and script.count("./b") == 2


def test_branchify_only_touches_expensive_conditionals():
    from absinthe_b200.emitter import branchify
    from absinthe_b200 import stencils as S
    cheap = S.diffusion_stencils()["ufli"]
    assert branchify(cheap) == cheap                      # the flux limiter stays a select
    adv = S.advection_stencils()
    out = branchify(adv["utev"])
    assert out.count("if (vatu(i,j,k) > 0.0)") == 1 and "?" not in out
    assert out.rstrip().endswith("auto res = uteu(i,j,k) + ( sel_0 );")
    top = branchify(adv["uteu"])
    assert top.count("if (uatu(i,j,k) > 0.0)") == 1 and "auto res = sel_0 ;" in top


def test_verify_emits_the_sequential_reference_variant(tmp_path):
    """VERIFY programs come with <name>_reference: one group per stencil, in compute_reference's order."""
    sequence = P.KERNELS["fastwaves"][1]()
    program = P.make_program("fastwaves", [sequence[:6], sequence[6:]], [(2, 4, 6), (1, 2, 3)],
                             variant="fastwaves-split", VERIFY=True)
    ref = G.sequential_reference(program)
    assert [g["GROUPS"][0]["STENCILS"] for g in ref["TILING"]["GROUPS"]] == [[name] for name in sequence]
    assert all((g["GROUPS"][0]["NX"], g["GROUPS"][0]["NY"], g["GROUPS"][0]["NZ"]) == (1, 8, 15)
               for g in ref["TILING"]["GROUPS"])
    assert ref["OUTPUTS"] == program["OUTPUTS"] and ref["VERIFY"] is False and ref["CUDA"] == {"LOAD": "ldg", "FMAD": False}
    assert ref["VARIANT"] == "fastwaves-split_reference"

    G.generate_code("template_tiling.cu", str(tmp_path / "fw.cu"), program)
    assert sorted(os.listdir(tmp_path)) == ["fw.cu", "fw.desc", "fw_reference.cu", "fw_reference.desc"]
    assert "# verify 1" in (tmp_path / "fw.desc").read_text().split("\n")
    desc = (tmp_path / "fw_reference.desc").read_text()
    assert "# verify" not in desc
    assert desc.count("\ngroup ") == len(sequence)
    # the same program analysed twice (a driver may pass an analysed dict) gives the same reference
    assert G.sequential_reference(program)["SEQUENCE"] == sequence
    # intermediates live in global memory and get the periodic update after their stencil
    for name in ("ppgk", "ppgc", "ppgu", "ppgv", "udc", "vdc"):
        assert "field %s tmp" % name in desc and "halo %s" % name in desc

    plain = P.make_program("advection", variant="plain")
    G.generate_makefile(str(tmp_path / "Makefile"), {"fw": program, "plain": plain})
    makefile = (tmp_path / "Makefile").read_text()
    assert "all: fw.cubin fw_reference.cubin plain.cubin\n" in makefile
    assert "fw_reference.cubin: fw_reference.cu\n" in makefile and "plain_reference" not in makefile


def test_node_grid_decomposes_along_j(tmp_path):
    """TILING.NY ranks (ref template_tiling.cpp:189-203): kernels are generated for one rank's slab, the descriptor
    records the node grid, run.sh launches one rank per GPU."""
    program = P.make_program("fastwaves", variant="fw2", domain=(64, 64, 20))
    program["TILING"]["NY"] = 2
    G.generate_code("template_tiling.cu", str(tmp_path / "fw2.cu"), program)
    desc = (tmp_path / "fw2.desc").read_text()
    assert "# nodes 1 2 1" in desc and "# global 64 64 20" in desc and "domain 64 32 20 3 3 3" in desc
    assert "constexpr int X = 64, Y = 32, Z = 20;" in (tmp_path / "fw2.cu").read_text()
    G.generate_script(str(tmp_path / "run.sh"), {"fw2": program}, 148, 1)
    script = (tmp_path / "run.sh").read_text()
    assert "torch.distributed.run --nnodes=1 --nproc-per-node 2" in script and "-m absinthe_b200.runner" in script
    bad = P.make_program("fastwaves", variant="fw3", domain=(64, 64, 20))
    bad["TILING"]["NX"] = 2
    with pytest.raises(AssertionError, match="along j only"):
        G.generate_code("template_tiling.cu", str(tmp_path / "fw3.cu"), bad)


def test_makefile_flags_are_per_target(tmp_path):
    strict = P.make_program("advection", variant="strict")
    relaxed = P.make_program("advection", variant="relaxed")
    relaxed["CUDA"] = {"FMAD": True}
    G.generate_makefile(str(tmp_path / "Makefile"), {"strict": strict, "relaxed": relaxed})
    rules = (tmp_path / "Makefile").read_text().split("\n\n")
    rule = {r.split(".cubin:")[0]: r for r in rules if ".cubin:" in r}
    assert "$(STRICT)" in rule["strict"] and "$(STRICT)" not in rule["relaxed"]


def test_verify_reference_uses_the_plain_kernels():
    program = P.make_program("fastwaves", variant="v", VERIFY=True)
    program["CUDA"] = {"INLINE": "all", "SHUFFLE": True, "FUSE_HALO": True, "FMAD": True}
    from absinthe_b200 import analysis as A
    A.analyse(program)
    ref = G.sequential_reference(program)
    assert ref["CUDA"] == {"LOAD": "ldg", "FMAD": False} and len(ref["TILING"]["GROUPS"]) == 9
