"""ORACLE -- TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never by the product package).

Renders ``oracle_template.cpp`` for one program, compiles it with g++ and wraps the
resulting shared object with ctypes.  The oracle is deliberately independent of
``absinthe_b200.analysis``: it derives inputs / temporaries with its own few lines
(a field that is read but never written is an input; every stencil result that is not
a program output is a full-size temporary), exactly what the reference's
``compute_reference`` allocates (ref template_tiling.cpp:162-170).
"""

import ctypes
import hashlib
import os
import re
import subprocess

import numpy as np
from jinja2 import Environment, FileSystemLoader

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")

STRICT_FLAGS = ["-std=c++11", "-O2", "-ffp-contract=off", "-fPIC", "-shared"]
# the flags the reference ships (ref stencil_generator.py:72-73), for the "as shipped" comparison
FAST_FLAGS = ["-std=c++11", "-O3", "-ffast-math", "-fPIC", "-shared"]

_ACCESS = re.compile(r"(\w+)\(\s*i+\s*(?:[+-]+\s*\d)?\s*,\s*j+\s*(?:[+-]+\s*\d)?\s*,\s*k+\s*(?:[+-]+\s*\d)?\s*\)")


def flatten_sequence(program):
    """Stencil names in execution order (groups flattened); falls back to SEQUENCE."""
    names = []
    for outer in program.get("TILING", {}).get("GROUPS", []):
        for inner in outer.get("GROUPS", []):
            for s in inner["STENCILS"]:
                names.append(s["NAME"] if isinstance(s, dict) else s)
    return names or list(program["SEQUENCE"])


def dataflow(program):
    sequence = flatten_sequence(program)
    written = set(sequence)
    read = []
    for name in sequence:
        for field in _ACCESS.findall(program["STENCILS"][name]):
            if field not in written and field not in read:
                read.append(field)
    outputs = sorted(program["OUTPUTS"])
    temps = [name for name in sequence if name not in outputs]
    return sorted(read), outputs, temps, sequence


class Oracle:
    """Compiled sequential oracle of one program."""

    def __init__(self, program, fast_math=False):
        inputs, outputs, temps, sequence = dataflow(program)
        self.inputs, self.outputs = inputs, outputs
        self.dims = tuple(program[k] for k in ("X", "Y", "Z", "HX", "HY", "HZ"))
        context = {k: program[k] for k in ("X", "Y", "Z", "HX", "HY", "HZ")}
        context.update(INPUTS=inputs, OUTPUTS=outputs, TEMPS=temps,
                       STENCILS=[{"NAME": n, "LAMBDA": program["STENCILS"][n]} for n in sequence])
        env = Environment(loader=FileSystemLoader(HERE))
        source = env.get_template("oracle_template.cpp").render(context)
        flags = FAST_FLAGS if fast_math else STRICT_FLAGS
        tag = hashlib.sha256((source + " ".join(flags)).encode()).hexdigest()[:20]
        os.makedirs(BUILD_DIR, exist_ok=True)
        self.path = os.path.join(BUILD_DIR, "oracle_%s.so" % tag)
        if not os.path.exists(self.path):
            cpp = os.path.join(BUILD_DIR, "oracle_%s.cpp" % tag)
            with open(cpp, "w") as handle:
                handle.write(source)
            tmp = self.path + ".%d.tmp" % os.getpid()
            subprocess.check_call(["g++"] + flags + [cpp, "-o", tmp])
            os.replace(tmp, self.path)
        self.lib = ctypes.CDLL(self.path)
        self.lib.oracle_field_name.restype = ctypes.c_char_p
        X, Y, Z, HX, HY, HZ = self.dims
        self.shape = (Z + 2 * HZ, Y + 2 * HY, X + 2 * HX)  # numpy order: k, j, i (i fastest)

    @staticmethod
    def _pointers(arrays):
        ptrs = (ctypes.c_void_p * len(arrays))()
        for n, a in enumerate(arrays):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
            ptrs[n] = a.ctypes.data
        return ptrs

    def empty_fields(self, count):
        return [np.zeros(self.shape, dtype=np.float64) for _ in range(count)]

    def fill_inputs(self, seed=0):
        """The reference's initialisation: srand(seed), rand() per cell, then periodic halos."""
        fields = self.empty_fields(len(self.inputs))
        self.lib.oracle_fill_inputs(self._pointers(fields), len(fields), ctypes.c_uint(seed))
        return fields

    def make_periodic(self, array):
        self.lib.oracle_make_periodic(ctypes.c_void_p(array.ctypes.data))
        return array

    def run(self, inputs):
        assert len(inputs) == len(self.inputs)
        for a in inputs:
            assert a.shape == self.shape, (a.shape, self.shape)
        outputs = self.empty_fields(len(self.outputs))
        self.lib.oracle_run(self._pointers(inputs), self._pointers(outputs))
        return outputs


def digest(array):
    """sha256 of the raw little-endian fp64 bytes -- the golden-vector currency."""
    return hashlib.sha256(np.ascontiguousarray(array).tobytes()).hexdigest()
