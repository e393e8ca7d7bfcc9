"""Host-side planning of the CUDA emitter: phases, inlining closure, shuffle lanes, streaming windows."""

import pytest

from absinthe_b200 import programs as P
from absinthe_b200.emitter import Plan, branchify

D = P.KERNELS["diffusion"][1]()
F = P.KERNELS["fastwaves"][1]()
FLUXES = [n for n in D if n.endswith("fli") or n.endswith("flj")]


def plan(kernel, groups, tiles, cuda, domain=(64, 64, 60)):
    program = P.make_program(kernel, groups, tiles, domain=domain)
    program["CUDA"] = dict(cuda)
    return Plan(program)


def test_generic_plan_keeps_every_intermediate_in_shared_memory():
    g = plan("diffusion", [D], [(2, 4, 15)], {}).groups[0]
    assert g.inlined == set() and sorted(g.in_smem) == sorted(n for n in D if not n.endswith("out"))
    # lap -> (fli, flj) -> out per field; equal level and box share a phase
    names = [ph["NAMES"] for ph in g.phase_list]
    assert names[0] == ["ulap", "vlap", "wlap", "pplap"] and names[-1] == ["uout", "vout", "wout", "ppout"]
    assert len(names) == 4          # laps | flis | fljs | outs (fli and flj have different boxes)


def test_inlining_fluxes_leaves_two_phases():
    g = plan("diffusion", [D], [(2, 4, 15)], {"INLINE": FLUXES}).groups[0]
    assert [ph["NAMES"] for ph in g.phase_list] == [["ulap", "vlap", "wlap", "pplap"],
                                                    ["uout", "vout", "wout", "ppout"]]
    passes = g.passes()
    assert [p["SYNC_BEFORE"] for p in passes] == [False, True]
    # the output phase reads the laplacian through the inlined fluxes at i-1 .. i+1
    assert g.closure["uout"]["ulap"] == ((-1, 1), (-1, 1), (0, 0))


def test_full_inlining_needs_no_shared_intermediates_and_no_barrier():
    g = plan("diffusion", [D], [(2, 4, 15)], {"INLINE": "all", "BLOCK": (4, 1)}).groups[0]
    assert g.in_smem == [] and len(g.phase_list) == 1
    p = g.passes()[0]
    assert not p["SYNC_BEFORE"] and len(p["STORES"]) == 16          # 4 outputs x 4 points per thread
    # radius-2 diamond of uin around a column of 4 points
    assert g.closure["uout"]["uin"] == ((-2, 2), (-2, 2), (0, 0))


def test_shuffle_lane_ranges():
    g = plan("diffusion", [D], [(2, 4, 15)], {"INLINE": "all", "SHUFFLE": True}).groups[0]
    p = g.passes()[0]
    assert p["LANES"] == (1, 30) and p["USEFUL"] == 30              # fli(i-1) and lap(i+1) come from neighbours
    body = "\n".join(p["BODY"])
    assert "__shfl_down_sync" in body and "__shfl_up_sync" in body
    a = plan("advection", None, None, {"INLINE": "all", "SHUFFLE": True}).groups[0]
    assert a.passes()[0]["LANES"] == (0, 31)                        # no intermediate is read at an i-offset


def test_outputs_read_later_are_kept_in_shared_memory_too():
    g = plan("fastwaves", [F], [(2, 4, 6)], {"INLINE": "all"}).groups[0]
    assert sorted(g.in_smem) == ["uout", "vout"] and g.inlined == {"ppgk", "ppgc", "ppgu", "ppgv", "udc", "vdc"}
    assert [ph["NAMES"] for ph in g.phase_list] == [["uout"], ["vout"], ["div"]]


def test_streaming_windows_and_rings():
    g = plan("fastwaves", [F], [(8, 128, 1)], {"STREAM_K": True, "INLINE": "all", "PREFETCH": 2},
             domain=(1024, 1024, 80)).groups[0]
    assert g.stream and g.B[2] == 80 and g.nsub[2] == 1 and g.kmin == -2
    # uout runs one plane ahead and reads ppuv at k-1 .. k+1 there: planes kk .. kk+2 are alive
    assert g.win["ppuv"] == (0, 2) and g.win["uin"] == (1, 1) and g.win["dzdx"] == (0, 0)
    rings = {p["NAME"]: p["RING"] for p in g.input_plan}
    assert rings["ppuv"] == 5 and rings["uin"] == 3 and rings["hhl"] == 4
    assert {t["NAME"]: t["DEPTH"] for t in g.temp_plan.values()} == {"uout": 4, "vout": 4}
    assert g.stages == 3                                            # prefetch + 1 steps in flight


def test_direct_inputs_are_centre_only():
    g = plan("diffusion", [D[:4]], [(1, 4, 15)], {"DIRECT": "auto", "INLINE": "all"}).groups[0]
    assert g.direct == {"mask"} and g.staged == ["uin"]


def test_subtile_respects_budget_and_block():
    for budget in (60 * 1024, 110 * 1024, 220 * 1024):
        g = plan("diffusion", [D], [(2, 64, 80)], {"INLINE": "all", "BLOCK": (4, 1), "SMEM_BUDGET": budget},
                 domain=(1024, 1024, 80)).groups[0]
        assert g.smem_bytes <= budget + 2048 and g.B[1] % 4 == 0


def test_branchify_leaves_nested_or_cheap_conditionals_alone():
    assert branchify("auto res = a > 0.0 ? b : c;") == "auto res = a > 0.0 ? b : c;"
    text = "auto res = x + (p > 0.0 ? p * (a+b+c+d+e) : p * (f+g+h+i+j));"
    out = branchify(text)
    assert out.startswith("double sel_0; if (p > 0.0)") and out.endswith("auto res = x + ( sel_0 );")


def test_impossible_shared_memory_plan_is_an_error():
    with pytest.raises(AssertionError):
        plan("fastwaves", [F], [(1, 1, 1)], {"SUBTILE": (256, 64, 60)})


def test_per_group_options_override_the_program_wide_ones():
    seq = P.KERNELS["fastwaves"][1]()
    program = P.make_program("fastwaves", [seq[:6], seq[6:]], [(2, 8, 4), (1, 4, 6)])
    program["CUDA"] = {"INLINE": "all", "BLOCK": (1, 2), "THREADS": 384,
                       "GROUPS": {"2": {"THREADS": 256, "BLOCK": (2, 1), "STAGES": 3}}}
    g1, g2 = Plan(P.clone(program)).groups
    assert (g1.threads, g1.block, g1.stages) == (384, (1, 1, 2), 2)
    assert (g2.threads, g2.block, g2.stages) == (256, (1, 2, 1), 3)
    text = Plan(P.clone(program)).descriptor_text()
    assert "group 1 kernel absinthe_fastwaves_g1 threads %d" % (384 + 32) in text
    assert "group 2 kernel absinthe_fastwaves_g2 threads %d" % (256 + 32) in text
    program["CUDA"]["GROUPS"] = {"1": {"FMAD": True}}
    with pytest.raises(AssertionError, match="not a per-group option"):
        Plan(P.clone(program))
