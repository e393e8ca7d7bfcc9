"""The oracle against the reference: committed digests of reference-computed inputs/outputs
(tests/golden/make_golden.py), and -- where the reference sources are mounted -- the live
compute_reference of the reference's own generated program."""

import json
import os

import numpy as np
import pytest

from absinthe_b200 import programs as P
from common import zoo
from oracle.reference_oracle import Oracle, digest

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_digests.json")))
CASES = zoo()


@pytest.mark.parametrize("name,program", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_reference_digests(name, program):
    gold = GOLDEN[name]
    oracle = Oracle(P.clone(program))
    assert sorted(oracle.inputs) == sorted(gold["input_order"])
    # initialise in the order the reference's main() constructed its arrays
    fields = oracle.empty_fields(len(gold["input_order"]))
    oracle.lib.oracle_fill_inputs(oracle._pointers(fields), len(fields), 0)
    by_name = dict(zip(gold["input_order"], fields))
    for field, want in gold["inputs"].items():
        assert digest(by_name[field]) == want, "input %s differs from the reference's rand() fill" % field
    outputs = oracle.run([by_name[n] for n in oracle.inputs])
    for field, array in zip(oracle.outputs, outputs):
        assert digest(array) == gold["outputs"][field], "output %s differs from compute_reference" % field
        assert float(np.abs(array).max()) == gold["abs_max"][field]


def test_outputs_are_periodic():
    oracle = Oracle(P.clone(CASES[0][1]))
    for array in oracle.run(oracle.fill_inputs()):
        before = array.copy()
        oracle.make_periodic(array)
        assert np.array_equal(before, array)


def test_fast_math_build_stays_within_normwise_tolerance():
    # SURVEY.md section 7 "Tolerance reality": the as-shipped -ffast-math build differs from strict IEEE
    # pointwise, but stays within 1e-12 of the field's magnitude.
    program = CASES[6][1]   # advection
    strict, fast = Oracle(P.clone(program)), Oracle(P.clone(program), fast_math=True)
    inputs = strict.fill_inputs()
    for a, b in zip(strict.run(inputs), fast.run(inputs)):
        assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources not mounted")
def test_oracle_against_live_reference():
    from oracle import build_ref
    program = CASES[8][1]   # fastwaves, two groups
    dump = build_ref.build_dump(program, "live_check")
    oracle = Oracle(P.clone(program))
    inputs = dump.fill_inputs()
    want = dump.compute(inputs)
    got = oracle.run([inputs[n] for n in oracle.inputs])
    for field, array in zip(oracle.outputs, got):
        assert array.tobytes() == want[field].tobytes()


TRAINING = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "training_digests.json")))


@pytest.mark.parametrize("name", sorted(TRAINING))
def test_oracle_reproduces_reference_digests_of_training_programs(name):
    """fitcache / fitddr stencils (ref fitcache.py:16-58, fitddr.py:16-124): values computed by the reference's own
    compute_reference for the training programs' stencils (tests/golden/make_golden.py: training_digests)."""
    family, variant, shape = name.split("-")
    program = P.training_program(family, variant, tuple(int(v) for v in shape.split("x")), cores=2, stride=20)
    gold = TRAINING[name]
    oracle = Oracle(P.clone(program))
    assert sorted(oracle.inputs) == sorted(gold["input_order"])
    fields = oracle.empty_fields(len(gold["input_order"]))
    oracle.lib.oracle_fill_inputs(oracle._pointers(fields), len(fields), 0)
    by_name = dict(zip(gold["input_order"], fields))
    for field, want in gold["inputs"].items():
        assert digest(by_name[field]) == want
    outputs = oracle.run([by_name[n] for n in oracle.inputs])
    assert sorted(oracle.outputs) == sorted(gold["outputs"])
    for field, array in zip(oracle.outputs, outputs):
        assert digest(array) == gold["outputs"][field], "output %s differs from compute_reference" % field
        assert float(np.abs(array).max()) == gold["abs_max"][field]
