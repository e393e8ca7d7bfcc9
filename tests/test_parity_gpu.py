"""GPU parity: every variant must reproduce the oracle bit-for-bit (fp64, -fmad=false),
on the full padded arrays (interior and periodic halos), through the C ABI."""

import pytest

from absinthe_b200 import programs as P
from common import assert_bit_exact, zoo

pytestmark = pytest.mark.gpu

CASES = zoo()


@pytest.mark.parametrize("name,program", CASES, ids=[c[0] for c in CASES])
def test_variant_matches_oracle(launcher_lib, name, program):
    assert_bit_exact(program)


@pytest.mark.parametrize("name,program", CASES[:3] + CASES[8:11], ids=[c[0] for c in CASES[:3] + CASES[8:11]])
def test_staged_loads_match_oracle(launcher_lib, name, program):
    program = P.clone(program)
    program["CUDA"] = {"LOAD": "ldg"}
    assert_bit_exact(program)


OPTIONS = [
    ("inline-all", {"INLINE": "all"}),
    ("inline-block2", {"INLINE": "all", "BLOCK": (2, 1)}),
    ("inline-block4x2-direct-halo", {"INLINE": "all", "BLOCK": (4, 2), "DIRECT": "auto", "FUSE_HALO": True}),
    ("block3-halo-ldg", {"BLOCK": (3, 1), "FUSE_HALO": True, "LOAD": "ldg", "DIRECT": "auto"}),
    ("halo-512", {"FUSE_HALO": True, "THREADS": 512, "SMEM_BUDGET": 200 * 1024}),
    ("shuffle", {"INLINE": "all", "SHUFFLE": True, "FUSE_HALO": True}),
    ("shuffle-block4x2", {"INLINE": "all", "SHUFFLE": True, "BLOCK": (4, 2), "DIRECT": "auto", "THREADS": 128}),
    ("branch", {"INLINE": "all", "BRANCH": True, "BLOCK": (2, 1), "FUSE_HALO": True}),
    ("cache-hints", {"INLINE": "all", "BLOCK": (2, 1), "FUSE_HALO": True, "STORE_CS": True, "L2_ONCE": True,
                     "L2_PREFETCH": 3}),
    ("uniform-warp", {"INLINE": "all", "SHUFFLE": True, "BLOCK": (2, 1), "FUSE_HALO": True, "UNIFORM_WARP": True}),
    ("uniform-warp-stream-k", {"STREAM_K": True, "INLINE": "all", "BLOCK": (2, 1), "FUSE_HALO": True,
                               "UNIFORM_WARP": True}),
    ("store-ptr", {"INLINE": "all", "SHUFFLE": True, "BLOCK": (4, 1), "FUSE_HALO": True, "UNIFORM_WARP": True,
                   "FAST_PATH": False, "STORE_PTR": True}),
    ("store-ptr-generic-cs", {"BLOCK": (2, 2), "STORE_PTR": True, "STORE_CS": True, "THREADS": 128}),
    ("store-ptr-stream-k", {"STREAM_K": True, "INLINE": "all", "BLOCK": (2, 1), "FUSE_HALO": True, "STORE_PTR": True}),
    ("stream-k", {"STREAM_K": True}),
    ("stream-k-inline", {"STREAM_K": True, "INLINE": "all", "BLOCK": (2, 1), "FUSE_HALO": True, "PREFETCH": 1}),
    ("poll-warp", {"INLINE": "all", "BLOCK": (2, 1), "FUSE_HALO": True, "POLL_WARP": True, "STAGES": 3}),
    ("poll-warp-generic", {"POLL_WARP": True, "THREADS": 128}),
    ("column", {"STREAM_K": "regs"}),
    ("column-auto", {"STREAM_K": "auto", "INLINE": "all", "BLOCK": (2, 1), "SHUFFLE": True, "FUSE_HALO": True}),
    ("column-merge", {"STREAM_K": "auto", "INLINE": "all", "BLOCK": (1, 1), "FUSE_HALO": True, "THREADS": 128,
                      "COLUMN": {"MERGE": True, "WARPS": (1, 3), "UNROLL": 1}}),
    ("column-cluster2", {"STREAM_K": "regs", "BLOCK": (1, 1), "FUSE_HALO": True, "THREADS": 128,
                         "COLUMN": {"MERGE": True, "WARPS": (1, 3), "UNROLL": 1, "CLUSTER": (2, 1), "SLACK": 2}}),
    ("column-cluster2x2-planes2", {"STREAM_K": "regs", "BLOCK": (2, 1), "FUSE_HALO": True,
                                   "COLUMN": {"MERGE": True, "WARPS": (1, 2), "PLANES": 2, "SLOTS": 2, "CLUSTER": (2, 2),
                                              "SLACK": 1}}),
    ("column-cluster1x8", {"STREAM_K": "regs", "BLOCK": (1, 1),
                           "COLUMN": {"MERGE": True, "WARPS": (2, 2), "UNROLL": 1, "CLUSTER": (1, 8)}}),
    ("column-rj2-planes2-halo", {"STREAM_K": "regs", "BLOCK": (2, 1), "COLUMN": {"WARPS": (1, 2), "PLANES": 2},
                                 "FUSE_HALO": True}),
    ("column-rj3-planes3", {"STREAM_K": "regs", "BLOCK": (3, 1),
                            "COLUMN": {"WARPS": (2, 1), "PLANES": 3, "SLOTS": 2, "UNROLL": 1}}),
    ("column-rj4-branch-halo", {"STREAM_K": "regs", "BLOCK": (4, 1), "COLUMN": {"WARPS": (1, 1)}, "BRANCH": True,
                                "FUSE_HALO": True, "STORE_CS": True}),
    ("stream-k-shuffle", {"STREAM_K": True, "INLINE": "all", "BLOCK": (3, 1), "SHUFFLE": True, "FUSE_HALO": True,
                          "PREFETCH": 3, "THREADS": 128}),
]
PARTIAL_INLINE = {"diffusion": ["ufli", "uflj", "vfli", "vflj", "wfli", "wflj", "ppfli", "ppflj"],
                  "advection": ["uatu", "vatu"], "fastwaves": ["ppgk", "ppgc", "udc"]}


@pytest.mark.parametrize("label,cuda", OPTIONS, ids=[o[0] for o in OPTIONS])
@pytest.mark.parametrize("name,program", CASES, ids=[c[0] for c in CASES])
def test_emitter_options_match_oracle(launcher_lib, name, program, label, cuda):
    program = P.clone(program)
    program["CUDA"] = dict(cuda)
    assert_bit_exact(program)


@pytest.mark.parametrize("name,program", CASES, ids=[c[0] for c in CASES])
def test_partial_inlining_matches_oracle(launcher_lib, name, program):
    program = P.clone(program)
    program["CUDA"] = {"INLINE": PARTIAL_INLINE[program["NAME"]], "BLOCK": (2, 2), "FUSE_HALO": True}
    assert_bit_exact(program)
    program["CUDA"]["SHUFFLE"] = True
    assert_bit_exact(program)
    program["CUDA"].update(STREAM_K=True, BLOCK=(2, 1))
    assert_bit_exact(program)


def test_cuda_graph_replay(launcher_lib):
    assert_bit_exact(CASES[0][1], graph=True)


@pytest.mark.parametrize("family,variant", [("fitcache", "PT12"), ("fitcache", "PT20"), ("fitddr", "IN0BD0"),
                                            ("fitddr", "IN2BD1"), ("fitddr", "IN3BD2")])
@pytest.mark.parametrize("stride", [20, 1, 7])
def test_training_protocol_matches_oracle_on_selected_tiles(launcher_lib, family, variant, stride):
    """Training mode evaluates tiles 0, 20, 40, ... only (ref template_training.cpp:277-278): those tiles
    must carry the oracle's values, every other cell of the outputs must stay untouched.  Other strides
    (``fitcache.py -s N``; 1 = every tile) select their own tiles."""
    import numpy as np
    import torch
    from absinthe_b200.emitter import tile_geometry
    from absinthe_b200.launcher import Variant
    from common import run_oracle
    cores = 3
    program = P.training_program(family, variant, (10, 3, 4), cores=cores, stride=stride)
    oracle, inputs, expected = run_oracle(program)
    v = Variant.from_program(P.clone(program), "template_training.cu")
    by_name = dict(zip(oracle.inputs, inputs))
    ins = [torch.from_numpy(by_name[n]).cuda() for n in v.inputs]
    outs = [torch.full(v.shape, -7.0, dtype=torch.float64, device="cuda") for _ in v.outputs]
    v.bind(ins, outs)
    total, halo = v.run(1, flush=False, training=True)
    assert halo == [0.0] and total[0] > 0
    counts = (2, 2, 5 * cores)
    dims = (v.X, v.Y, v.Z)
    geo = [tile_geometry(s, n) for s, n in zip(dims, counts)]
    mask = np.zeros(v.shape, dtype=bool)
    for tile in range(0, counts[0] * counts[1] * counts[2], stride):
        t = (tile % counts[0], (tile // counts[0]) % counts[1], tile // (counts[0] * counts[1]))
        lo = [max(0, c * g[0] + g[1]) for c, g in zip(t, geo)]
        hi = [min(s, (c + 1) * g[0] + g[1]) for c, g, s in zip(t, geo, dims)]
        mask[3 + lo[2]:3 + hi[2], 3 + lo[1]:3 + hi[1], 3 + lo[0]:3 + hi[0]] = True
    assert mask.sum() == len(range(0, 20 * cores, stride)) * 10 * 3 * 4
    for name, want in zip(oracle.outputs, expected):
        got = outs[v.outputs.index(name)].cpu().numpy()
        assert np.array_equal(got[mask], want[mask]), name
        assert np.all(got[~mask] == -7.0), name
    v.close()


def _shipped_cases(per_kernel=5):
    """Fusion groupings of variants the reference ships estimates for (names like
    ``diffusion-0-0-1-1-...-mz``): the grouping is in the name, the tile counts are not pinned
    (SURVEY.md section 8c), so deterministic pseudo-random counts are used."""
    import json
    import os
    import random
    names = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "shipped_variants.json")))
    rng = random.Random(2019)
    cases = []
    for kernel, variants in sorted(names.items()):
        seq = P.KERNELS[kernel][1]()
        grouped = [v for v in variants if v.startswith(kernel + "-")]
        for name in rng.sample(grouped, per_kernel):
            digits = [int(t) for t in name.split("-")[1:1 + len(seq)]]
            groups = P.split_sequence(seq, digits)
            tiles = [(rng.choice([1, 2, 3, 5]), rng.choice([1, 2, 4, 7]), rng.choice([1, 2, 3])) for _ in groups]
            cases.append((name, P.make_program(kernel, groups, tiles, domain=(40, 28, 12), variant=name)))
    return cases


SHIPPED = _shipped_cases()


@pytest.mark.parametrize("name,program", SHIPPED, ids=[c[0] for c in SHIPPED])
def test_shipped_variant_groupings_match_oracle(launcher_lib, name, program):
    assert_bit_exact(program)
    tuned = P.clone(program)
    tuned["CUDA"] = {"INLINE": "all", "BLOCK": (2, 1), "SHUFFLE": True, "FUSE_HALO": True}
    assert_bit_exact(tuned)


def test_per_group_options_match_oracle(launcher_lib):
    """Every group kernel of a variant may have its own launch shape / staging options."""
    seq = P.KERNELS["fastwaves"][1]()
    program = P.make_program("fastwaves", [seq[:4], seq[4:6], seq[6:]], [(2, 4, 3), (1, 8, 2), (4, 2, 6)],
                             domain=(72, 40, 18))
    program["CUDA"] = {"INLINE": "all", "FUSE_HALO": True, "BLOCK": (1, 2), "THREADS": 384,
                       "GROUPS": {"2": {"THREADS": 128, "BLOCK": (2, 1), "SHUFFLE": True, "STAGES": 3},
                                  "3": {"LOAD": "ldg", "BLOCK": (1, 1), "INLINE": "none"}}}
    assert_bit_exact(program)


@pytest.mark.parametrize("name,program", [CASES[1], CASES[5], CASES[10]], ids=[CASES[n][0] for n in (1, 5, 10)])
@pytest.mark.parametrize("label,cuda", [("classic", {"INLINE": "all", "BLOCK": (2, 1), "FUSE_HALO": True}),
                                        ("column", {"STREAM_K": "regs", "BLOCK": (2, 1), "FUSE_HALO": True})])
def test_fma_build_stays_within_the_stated_tolerance(launcher_lib, name, program, label, cuda):
    """``CUDA.FMAD`` (nvcc contracts a*b+c into fma; north_star allows <= 1e-12 in fp64): not bit-exact any more, but
    within 1e-12 of the strict oracle relative to the field's magnitude (raw rand() inputs cancel catastrophically
    pointwise, like the reference's own -ffast-math build, see tests/test_oracle.py)."""
    import numpy as np
    from common import run_gpu, run_oracle
    program = P.clone(program)
    program["CUDA"] = dict(cuda, FMAD=True)
    oracle, inputs, expected = run_oracle(program)
    got = run_gpu(program, dict(zip(oracle.inputs, inputs)))
    differs = 0
    for out, want in zip(oracle.outputs, expected):
        scale = np.abs(want).max()
        assert np.abs(got[out] - want).max() <= 1e-12 * scale, out
        differs += int((got[out] != want).sum())
    assert differs > 0                  # the option does something: a build that stayed bit-exact would not be FMA
