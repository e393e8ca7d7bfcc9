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


def test_cuda_graph_replay(launcher_lib):
    assert_bit_exact(CASES[0][1], graph=True)
