"""The column planner (absinthe_b200/column.py): k-leads, thread-local rows, register carries and
shuffle lanes of every group of the program zoo, verified by replaying the step program on symbolic
tags (ColumnPlan.simulate) -- no GPU needed.  The GPU parity cases are in test_parity_gpu.py."""

import subprocess

import pytest

from absinthe_b200 import programs as P
from absinthe_b200.column import ColumnPlan
from absinthe_b200.emitter import Plan
from absinthe_b200.generator import nvcc_command, render
from common import zoo

F = P.KERNELS["fastwaves"][1]()
SHAPES = [
    ("default", {}),
    ("rj2-planes2", {"BLOCK": (2, 1), "COLUMN": {"WARPS": (1, 2), "PLANES": 2}, "FUSE_HALO": True}),
    ("rj3-planes3", {"BLOCK": (3, 1), "COLUMN": {"WARPS": (2, 1), "PLANES": 3, "SLOTS": 2, "UNROLL": 1}}),
    ("rj4", {"BLOCK": (4, 1), "COLUMN": {"WARPS": (1, 1)}, "BRANCH": True}),
]


def column_groups(program, cuda):
    q = P.clone(program)
    q["CUDA"] = dict(cuda, STREAM_K="regs")
    return Plan(q).groups


@pytest.mark.parametrize("label,cuda", SHAPES, ids=[s[0] for s in SHAPES])
@pytest.mark.parametrize("name,program", zoo(), ids=[c[0] for c in zoo()])
def test_step_program_holds_the_right_cells(name, program, label, cuda):
    groups = column_groups(program, cuda)
    if program["X"] % 2:
        assert not any(g.column for g in groups)      # odd rows cannot be TMA sources: classic kernel
        return
    for g in groups:
        assert isinstance(g, ColumnPlan)
        assert g.simulate() == len(g.outputs) * g.rj * g.useful * (4 - g.kmin)
        assert g.smem_bytes <= 227 * 1024 and g.kmin <= 0


def test_fastwaves_schedule():
    """ref fastwaves.py:16-48: ppgc reads ppgk(k+1), udc/vdc read uout/vout(k-1), div reads udc/vdc(k+1)."""
    g = column_groups(P.make_program("fastwaves", [F], [(2, 4, 1)]), {"BLOCK": (2, 1)})[0]
    assert g.lead == {"ppgk": 2, "ppgc": 1, "ppgu": 1, "ppgv": 1, "uout": 1, "vout": 1, "udc": 1, "vdc": 1, "div": 0}
    assert g.qtop["ppuv"] == 2 and g.qtop["rho"] == 1 and g.qtop["dzdx"] == 0
    assert g.kmin == -4                                   # ppuv(k-2) enters at its newest plane k+2
    # ppgv(j) needs ppgc(j+1); div(j) needs vout(j-1): the v-chain is evaluated on one extra row below
    assert g.rows["vout"] == [-1, 0, 1] and g.rows["ppgc"] == [-1, 0, 1, 2] and g.rows["uout"] == [0, 1]
    assert (g.first, g.last) == (1, 30)                   # uout(i-1) and ppgc(i+1) come from the neighbour lanes
    # every stencil is evaluated once per row and step: nothing is recomputed along k
    evals = [op for op in g.ops if op[0] == "eval"]
    assert len(evals) == sum(len(g.rows[n]) for n in g.names)
    body = "\n".join(g._body())
    assert "__shfl_up_sync" in body and "__shfl_down_sync" in body and "c_uout_0_0_q0" in g.carried


def test_groups_without_k_dependence_carry_nothing():
    d = P.KERNELS["diffusion"][1]()
    g = column_groups(P.make_program("diffusion", [d], [(2, 2, 15)]), {"BLOCK": (2, 1)})[0]
    assert set(g.lead.values()) == {0} and g.carried == [] and g.kmin == 0
    assert g.rows["ulap"] == [-1, 0, 1, 2] and g.rows["uflj"] == [-1, 0, 1]   # uout reads uflj(j-1), uflj reads ulap(j+1)


def test_unrelated_options_fall_back_to_the_classic_kernel():
    q = P.make_program("fastwaves", [F], [(2, 4, 6)])
    q["CUDA"] = {"STREAM_K": "regs", "LOAD": "ldg"}
    assert not Plan(q).groups[0].column


@pytest.mark.parametrize("cuda", [SHAPES[1][1], SHAPES[2][1]], ids=["rj2", "rj3"])
def test_column_kernel_compiles_for_sm_100a(tmp_path, cuda):
    q = P.make_program("fastwaves", [F], [(2, 4, 2)])
    q["CUDA"] = dict(cuda, STREAM_K="regs")
    source, descriptor, flags = render("template_tiling.cu", q)
    assert "absinthe_fastwaves_g1" in source and "column kernel" in source
    assert "kernel absinthe_fastwaves_g1 threads %d" % (32 * (2 + 1)) in descriptor
    path = tmp_path / "column.cu"
    path.write_text(source)
    proc = subprocess.run(nvcc_command(flags, str(path), str(tmp_path / "column.cubin")), capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    sass = subprocess.run(["cuobjdump", "-sass", str(tmp_path / "column.cubin")], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass and "SHFL" in sass and "BAR.SYNC" in sass
    assert sass.count("BAR.SYNC") == 1                    # only the start-up barrier: no CTA barrier in the k-loop


def test_auto_picks_column_kernels_for_k_dependent_groups_only():
    q = P.make_program("fastwaves", [F[:4], F[4:6], F[6:]], [(2, 4, 3), (1, 8, 2), (4, 2, 6)])
    q["CUDA"] = {"STREAM_K": "auto", "INLINE": "all"}
    groups = Plan(q).groups
    assert [g.column for g in groups] == [True, False, True]      # ppgc reads ppgk(k+1); div reads udc(k+1)
    assert not groups[1].stream and groups[1].inlined == set()    # the classic kernel, not the plane-ring one
    d = P.KERNELS["diffusion"][1]()
    q = P.make_program("diffusion", [d], [(2, 4, 3)])
    q["CUDA"] = {"STREAM_K": "auto"}
    assert [g.column for g in Plan(q).groups] == [False]


def test_merge_walks_columns_of_the_kernels_own_shape():
    q = P.make_program("fastwaves", [F], [(9, 16, 6)], domain=(128, 64, 60))
    q["CUDA"] = {"STREAM_K": "regs", "BLOCK": (2, 1), "COLUMN": {"WARPS": (2, 3), "MERGE": True}}
    g = Plan(q).groups[0]
    assert g.N == (3, 11, 1) and g.T == (43, 6, 60) and g.nsub == (1, 1, 1)     # 2 strips of 30 columns, 3 x 2 rows
    assert g.simulate() == len(g.outputs) * g.rj * g.useful * (4 - g.kmin)
    assert "tiles 33 " in Plan(P.clone(q)).descriptor_text()


def test_cluster_kernels_walk_super_tiles(tmp_path):
    """COLUMN.CLUSTER: CX x CY neighbouring tiles per thread-block cluster; the producers pace each other through
    remote mbarrier arrives.  Only for merged columns of full-domain programs."""
    q = P.make_program("fastwaves", [F], [(9, 16, 6)], domain=(128, 64, 60))
    q["CUDA"] = {"STREAM_K": "regs", "COLUMN": {"WARPS": (2, 3), "MERGE": True, "CLUSTER": (2, 2), "SLACK": 3}}
    g = Plan(q).groups[0]
    assert g.N == (3, 22, 1) and g.cluster == (2, 2) and g.super == (2, 11) and g.item_chunks == 60 - g.kmin
    assert g.smem_bytes - g.tick_offset >= 8 * (2 * 3 + 1)
    source, descriptor, flags = render("template_tiling.cu", q)
    assert "__cluster_dims__(4, 1, 1)" in source and "mbar_arrive_remote(&tick[" in source
    assert " cluster 4" in descriptor
    path = tmp_path / "cluster.cu"
    path.write_text(source)
    proc = subprocess.run(nvcc_command(flags, str(path), str(tmp_path / "cluster.cubin")), capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    sass = subprocess.run(["cuobjdump", "-sass", str(tmp_path / "cluster.cubin")], capture_output=True, text=True).stdout
    # cluster barrier at start-up, remote arrives on the peers' tick barriers
    assert "UCGABAR_ARV" in sass and "UCGABAR_WAIT" in sass and "SYNCS.ARRIVE.TRANS64.RED" in sass
    # without MERGE (sub-tiles / k-tiles differ per work item) and for the training loop: plain column kernel
    q["CUDA"]["COLUMN"] = {"WARPS": (2, 3), "CLUSTER": (2, 2)}
    g = Plan(P.clone(q)).groups[0]
    assert g.cluster == (1, 1) and " cluster " not in Plan(P.clone(q)).descriptor_text()
    t = P.training_program("fitddr", "IN2BD1", (10, 3, 4), cores=3)
    t["CUDA"] = {"STREAM_K": "regs", "COLUMN": {"MERGE": True, "CLUSTER": (2, 1)}}
    assert all(getattr(g, "cluster", (1, 1)) == (1, 1) for g in Plan(t).groups)
