"""Shared helpers of the parity tests: program zoo + oracle/GPU comparison."""

import numpy as np

from absinthe_b200 import programs as P


def seq(kernel):
    return P.KERNELS[kernel][1]()


def zoo():
    """(id, program) pairs: small domains the oracle finishes in well under a second."""
    d, a, f = seq("diffusion"), seq("advection"), seq("fastwaves")
    cases = [
        ("diffusion-4groups", P.make_program("diffusion", P.split_sequence(d, [0] * 4 + [1] * 4 + [2] * 4 + [3] * 4),
                                             [(1, 4, 15)] * 4)),
        ("diffusion-fused", P.make_program("diffusion", [d], [(2, 2, 15)])),
        ("diffusion-unfused", P.make_program("diffusion", [[s] for s in d], [(1, 2, 5)] * 16, domain=(32, 24, 10))),
        ("diffusion-ragged", P.make_program("diffusion", P.split_sequence(d, [0] * 7 + [1] * 9), [(3, 5, 7), (7, 3, 11)],
                                            domain=(50, 37, 23))),
        ("diffusion-odd-x", P.make_program("diffusion", [d], [(2, 3, 2)], domain=(33, 21, 6))),
        ("advection-hand", P.make_program("advection", [a], [(1, 1, 60)])),
        ("advection-auto", P.make_program("advection", [a], [(1, 4, 5)])),
        ("advection-split", P.make_program("advection", P.split_sequence(a, [0, 0, 1, 1, 2, 2, 3, 3]),
                                           [(2, 2, 2), (1, 3, 4), (4, 1, 1), (1, 1, 9)], domain=(40, 28, 9))),
        ("fastwaves-hand", P.make_program("fastwaves", [f[:6], f[6:]], [(1, 8, 8), (1, 8, 8)])),
        ("fastwaves-auto", P.make_program("fastwaves", [f[:6], f[6:]], [(1, 4, 5), (1, 2, 5)])),
        ("fastwaves-fused", P.make_program("fastwaves", [f], [(2, 4, 6)])),
        ("fastwaves-unfused", P.make_program("fastwaves", [[s] for s in f], [(1, 2, 3)] * 9, domain=(24, 18, 12))),
        ("fastwaves-ragged", P.make_program("fastwaves", P.split_sequence(f, [0, 0, 1, 1, 1, 2, 2, 3, 3]),
                                            [(3, 2, 5), (2, 5, 3), (1, 7, 2), (5, 1, 4)], domain=(46, 31, 17))),
    ]
    return cases


def run_oracle(program, seed=0):
    """Reference-style inputs and the oracle's outputs: (oracle, inputs[list], outputs[list])."""
    from oracle.reference_oracle import Oracle
    oracle = Oracle(P.clone(program))
    inputs = oracle.fill_inputs(seed)
    return oracle, inputs, oracle.run(inputs)


def run_gpu(program, inputs_by_name, template="template_tiling.cu", graph=False):
    """One application of the variant on cuda:0; returns dict name -> numpy array."""
    import torch
    from absinthe_b200.launcher import Variant
    variant = Variant.from_program(P.clone(program), template)
    ins = [torch.from_numpy(np.ascontiguousarray(inputs_by_name[n])).cuda() for n in variant.inputs]
    outs = [variant.empty_field() for _ in variant.outputs]
    variant.bind(ins, outs)
    variant.apply(graph=graph)
    torch.cuda.synchronize()
    result = {n: t.cpu().numpy() for n, t in zip(variant.outputs, outs)}
    variant.close()
    return result


def assert_bit_exact(program, **kw):
    oracle, inputs, expected = run_oracle(program)
    got = run_gpu(program, dict(zip(oracle.inputs, inputs)), **kw)
    assert sorted(got) == sorted(oracle.outputs)
    for name, want in zip(oracle.outputs, expected):
        have = got[name]
        if have.tobytes() != want.tobytes():
            bad = np.argwhere(have != want)
            raise AssertionError("%s: %d of %d cells differ, first at (k,j,i)=%s: %r vs %r" % (
                name, len(bad), want.size, tuple(bad[0]), have[tuple(bad[0])], want[tuple(bad[0])]))
