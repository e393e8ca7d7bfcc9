"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Recipe that builds the *real* reference for one program.

Runs only where ``/root/reference`` exists (the development container).  It drives the
reference's own ``stencil_generator.generate_code`` (ref stencil_generator.py:41-59) from the
sources where they lie -- nothing is copied into the repository -- and compiles the rendered
program with g++.  All outputs go to ``oracle/_ref/`` (git-ignored, but shipped to the GPU box):

* ``<name>.cpp``        the rendered reference program (reference text: never committed)
* ``<name>``            the reference OpenMP benchmark binary, reference flags
                        (ref stencil_generator.py:72-73) -- the ``cpu_baseline`` of bench.py
* ``<name>_dump.so``    the same translation unit with ``main`` renamed, plus a small wrapper
                        (written here) exporting the reference's ``compute_reference``
                        (ref template_tiling.cpp:162-177) and its rand() input initialisation
                        (ref :209-221) through a C ABI, so the oracle restatement can be pinned
                        bit-for-bit against reference-computed field values.
"""

import ctypes
import os
import subprocess
import sys
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("ABSINTHE_REFERENCE", "/root/reference")

REFERENCE_FLAGS = ["-std=c++11", "-O3", "-ffast-math", "-fopenmp", "-DNDEBUG"]
STRICT_FLAGS = ["-std=c++11", "-O2", "-ffp-contract=off", "-fopenmp", "-DNDEBUG"]


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "stencil_generator.py"))


def render_reference(program, name, template="template_tiling.cpp"):
    """Render ``program`` with the reference generator; returns (cpp path, analysed dict)."""
    assert available(), "reference sources not present"
    os.makedirs(REF_DIR, exist_ok=True)
    program = deepcopy(program)
    # the reference expects plain stencil-name lists
    for outer in program["TILING"]["GROUPS"]:
        for inner in outer["GROUPS"]:
            inner["STENCILS"] = [s["NAME"] if isinstance(s, dict) else s for s in inner["STENCILS"]]
    cpp = os.path.join(REF_DIR, name + ".cpp")
    cwd = os.getcwd()
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        os.chdir(REFERENCE_ROOT)  # the reference loads its template from the cwd
        import stencil_generator as reference_generator
        reference_generator.generate_code(template, cpp, program)
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_ROOT)
    return cpp, program


def build_binary(program, name, template="template_tiling.cpp", flags=None):
    """The reference benchmark executable for ``program`` (timed as the CPU baseline)."""
    cpp, analysed = render_reference(program, name, template)
    exe = os.path.join(REF_DIR, name)
    subprocess.check_call(["g++"] + (flags or REFERENCE_FLAGS) + [cpp, "-o", exe])
    return exe, analysed


_WRAPPER = r"""
#define main absinthe_reference_main
#include "%(cpp)s"
#undef main
#include <cstring>
namespace {
constexpr long CELLS = (long)(X + 2 * HX) * (Y + 2 * HY) * (Z + 2 * HZ);
void load(array_3d& dst, const double* src) { std::memcpy(&dst(0, 0, 0), src, sizeof(double) * CELLS); }
void store(double* dst, array_3d& src) { std::memcpy(dst, &src(0, 0, 0), sizeof(double) * CELLS); }
}
extern "C" void ref_fill_inputs(double* const* inputs) {
    std::srand(0);
%(fill)s
}
extern "C" void ref_compute(const double* const* inputs, double* const* outputs) {
%(body)s
}
"""


def build_dump(program, name, flags=None):
    """Shared object exposing the reference's compute_reference for ``program``."""
    cpp, analysed = render_reference(program, name)
    ins, outs = analysed["TILING"]["INPUTS"], analysed["TILING"]["OUTPUTS"]
    fill = []
    for n, field in enumerate(ins):  # construction order == the order of the reference's main()
        fill.append("    array_3d _%s;" % field)
    for n, field in enumerate(ins):
        fill.append("    make_periodic(_%s); store(inputs[%d], _%s);" % (field, n, field))
    body = []
    for n, field in enumerate(ins):
        body.append("    array_3d _%s(0.0); load(_%s, inputs[%d]);" % (field, field, n))
    for field in outs:
        body.append("    array_3d _%s(0.0);" % field)
    body.append("    compute_reference(%s);" % ", ".join("_" + f for f in ins + outs))
    for n, field in enumerate(outs):
        body.append("    store(outputs[%d], _%s);" % (n, field))
    wrapper = os.path.join(REF_DIR, name + "_dump.cpp")
    with open(wrapper, "w") as handle:
        handle.write(_WRAPPER % {"cpp": cpp, "fill": "\n".join(fill), "body": "\n".join(body)})
    lib = os.path.join(REF_DIR, name + "_dump.so")
    subprocess.check_call(["g++"] + (flags or STRICT_FLAGS) + ["-fPIC", "-shared", "-Wno-return-type", wrapper, "-o", lib])
    return ReferenceDump(lib, analysed)


class ReferenceDump:
    """ctypes view of ``<name>_dump.so`` (field order is the reference's own, hash-seed dependent)."""

    def __init__(self, path, analysed):
        self.lib = ctypes.CDLL(path)
        self.inputs = list(analysed["TILING"]["INPUTS"])
        self.outputs = list(analysed["TILING"]["OUTPUTS"])
        p = analysed
        self.shape = (p["Z"] + 2 * p["HZ"], p["Y"] + 2 * p["HY"], p["X"] + 2 * p["HX"])

    @staticmethod
    def _pointers(arrays):
        ptrs = (ctypes.c_void_p * len(arrays))()
        for n, a in enumerate(arrays):
            ptrs[n] = a.ctypes.data
        return ptrs

    def fill_inputs(self):
        """name -> array, filled exactly as the reference's main() fills its inputs."""
        fields = [np.zeros(self.shape) for _ in self.inputs]
        self.lib.ref_fill_inputs(self._pointers(fields))
        return dict(zip(self.inputs, fields))

    def compute(self, inputs):
        """inputs: name -> periodic array; returns name -> array from compute_reference."""
        ordered = [np.ascontiguousarray(inputs[n]) for n in self.inputs]
        outs = [np.zeros(self.shape) for _ in self.outputs]
        self.lib.ref_compute(self._pointers(ordered), self._pointers(outs))
        return dict(zip(self.outputs, outs))
