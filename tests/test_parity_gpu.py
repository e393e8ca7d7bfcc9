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


def test_cuda_graph_replay(launcher_lib):
    assert_bit_exact(CASES[0][1], graph=True)
