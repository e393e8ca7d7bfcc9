#!/usr/bin/env python
"""Where does a variant differ from the oracle?  debug_diff.py <zoo case> '<python dict of CUDA options>'
Prints, per output, how many interior / halo cells differ and the padded k, j, i indices they sit on."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ranges(values):
    out, start, prev = [], None, None
    for v in sorted(set(int(x) for x in values)):
        if start is None:
            start = prev = v
        elif v == prev + 1:
            prev = v
        else:
            out.append((start, prev))
            start = prev = v
    if start is not None:
        out.append((start, prev))
    return " ".join("%d" % a if a == b else "%d-%d" % (a, b) for a, b in out)


def main():
    from absinthe_b200 import programs as P
    from common import run_gpu, run_oracle, zoo
    program = P.clone(dict(zoo())[sys.argv[1]])
    program["CUDA"] = eval(sys.argv[2])
    oracle, inputs, expected = run_oracle(program)
    got = run_gpu(program, dict(zip(oracle.inputs, inputs)))
    H = 3
    for name, want in zip(oracle.outputs, expected):
        bad = got[name] != want
        inner = bad[H:-H, H:-H, H:-H].sum()
        print("%s: %d cells differ (%d interior)" % (name, bad.sum(), inner))
        if bad.any():
            k, j, i = np.nonzero(bad)
            print("   k:", ranges(k), "\n   j:", ranges(j), "\n   i:", ranges(i))
            if inner:
                k, j, i = np.nonzero(bad[H:-H, H:-H, H:-H])
                print("   interior k:", ranges(k), "\n   interior j:", ranges(j), "\n   interior i:", ranges(i))


if __name__ == "__main__":
    main()
