"""Program-dict builders for the kernels of the reference's experiment drivers.

A *program* is the plain dict the reference's drivers assemble
(ref diffusion.py:67-86, advection.py:57-76, fastwaves.py:51-70,
fitcache.py:61-85, fitddr.py:127-156) and hand to ``generate_code``.  The
schema is kept verbatim; only the machine-model numbers
differ because they describe a B200 instead of the authors' 4-core CPU
(see ``fit.py``).
"""

from copy import deepcopy

from . import stencils as S

# GPU re-parameterisation of MACHINE: CORES <-> SMs,
# CAPACITY <-> shared memory one CTA (= one tile) may hold.
B200_SMS = 148
B200_SMEM_PER_CTA = 227 * 1024

# Analytical-model parameters.  MEMORY/CACHE start from the reference's CPU
# fit (ref diffusion.py:72-73) and are replaced by the GPU fit that
# ``fitcache.py`` / ``fitddr.py`` + ``absinthe_b200.fit`` produce.
CPU_MEMORY = {"RW BODY": -2.23e-7, "ST BODY": 5.71e-7, "RW PEEL": -1.25e-6, "ST PEEL": 5.25e-6}
CPU_CACHE = {"BODY": 9.44e-8, "PEEL": 9.95e-7}

# GPU fit of the training sweeps, round 2 (profiles/r02_fit.json): `fitcache.py -g -s 1` / `fitddr.py -g -s 1` on the
# B200 -- every tile of a training program runs (20 per SM), so the fit sees throughput, not one launch latency --
# fitted with absinthe_b200/fit.py (LAD, --tiles-per-core 20 --intercept).  CACHE: R2 0.957 on the 309 fitcache
# programs with >= 12 fetches.  MEMORY: the IN1..3 BD0 programs (R2 0.94, all coefficients positive without
# constraint); with boundary depth >= 1 the 27 haloed input streams of a fitddr program no longer fit one
# shared-memory stage, the kernels fall back to sub-tiles of a few dozen points and leave the linear regime
# (R2 0.31 over all programs).  Round 1 (one wave of tiles = launch latency) had
# RW 0 / ST BODY 4.91e-9 / ST PEEL 2.75e-8 and CACHE 1.17e-10 / 2.33e-10.
B200_MEMORY = {"RW BODY": 7.31e-10, "ST BODY": 1.258e-9, "RW PEEL": 2.2081e-8, "ST PEEL": 5.849e-9}
B200_CACHE = {"BODY": 7.016e-11, "PEEL": 3.950e-11}

KERNELS = {
    "diffusion": (S.diffusion_stencils, S.diffusion_sequence, S.DIFFUSION_OUTPUTS, S.DIFFUSION_CONSTANTS),
    "advection": (S.advection_stencils, S.advection_sequence, S.ADVECTION_OUTPUTS, S.ADVECTION_CONSTANTS),
    "fastwaves": (S.fastwaves_stencils, S.fastwaves_sequence, S.FASTWAVES_OUTPUTS, S.FASTWAVES_CONSTANTS),
}


def base_program(kernel, machine=None, memory=None, cache=None):
    """The driver-level PROGRAM dict of ``kernel`` (no STENCILS/TILING yet)."""
    _, _, outputs, constants = KERNELS[kernel]
    return {
        "NAME": kernel,
        "CONSTANTS": list(constants),
        "OUTPUTS": list(outputs),
        "MACHINE": dict(machine or {"CORES": B200_SMS, "CAPACITY": B200_SMEM_PER_CTA}),
        "MEMORY": dict(memory or B200_MEMORY),
        "CACHE": dict(cache or B200_CACHE),
        "OVERLAP": 1.0,
        "SLACK": {"SIZE": 0.02, "CORES": 0.05},
        "CONSTRAINTS": {},
        "X": 64, "Y": 64, "Z": 60,
        "HX": 3, "HY": 3, "HZ": 3,
        "RUNS": 64,
        "VERIFY": False,
        "FLUSH": True,
    }


def tiling_from_groups(groups, tiles, nodes=(1, 1, 1)):
    """TILING dict from a list of stencil-name lists and matching (NX, NY, NZ) tile counts."""
    assert len(groups) == len(tiles)
    return {
        "NX": nodes[0], "NY": nodes[1], "NZ": nodes[2],
        "GROUPS": [{"GROUPS": [{"STENCILS": list(names), "NX": t[0], "NY": t[1], "NZ": t[2]}]}
                   for names, t in zip(groups, tiles)],
    }


def split_sequence(sequence, assignment):
    """Cut ``sequence`` into consecutive groups by a per-stencil group index list."""
    groups = []
    last = None
    for name, g in zip(sequence, assignment):
        if g != last:
            groups.append([])
            last = g
        groups[-1].append(name)
    return groups


def make_program(kernel, groups=None, tiles=None, domain=None, variant=None, outputs=None, **extra):
    """A ready-to-generate program: full sequence fused into ``groups`` with ``tiles`` counts.

    ``groups`` defaults to one group holding the whole sequence, ``tiles`` to one
    tile per group.  ``outputs`` restricts the program outputs (the auto-tuning
    sub-programs of the drivers do that, ref diffusion.py:108,121).
    """
    stencil_fn, sequence_fn, _, _ = KERNELS[kernel]
    program = base_program(kernel)
    sequence = sequence_fn()
    if groups is None:
        groups = [sequence]
    if tiles is None:
        tiles = [(1, 1, 1)] * len(groups)
    used = [name for g in groups for name in g]
    table = stencil_fn()
    program["STENCILS"] = {name: table[name] for name in table if name in used}
    program["SEQUENCE"] = used
    if outputs is not None:
        program["OUTPUTS"] = list(outputs)
    if domain is not None:
        program["X"], program["Y"], program["Z"] = domain
    program["VARIANT"] = variant or kernel
    program["TILING"] = tiling_from_groups(groups, tiles)
    program.update(extra)
    return program


# ---------------------------------------------------------------------------------------
# Training programs (ref fitcache.py:61-85,116-172; fitddr.py:127-156,187-327)
# ---------------------------------------------------------------------------------------
TRAINING_X = (10, 20, 30, 50, 80)
TRAINING_YZ = (1, 2, 3, 5, 8, 13, 21, 34, 55)


def training_tile_shapes():
    """The 103 (xlen, ylen, zlen) tile shapes with 500..2000 points (ref fitcache.py:117-121)."""
    return [(x, y, z) for x in TRAINING_X for y in TRAINING_YZ for z in TRAINING_YZ
            if 500 <= x * y * z <= 2000]


def training_program(family, variant, shape, cores=B200_SMS, stride=20):
    """One training program: 2 x 2 x (5*cores) tiles of ``shape`` points, all nine stencils fused.

    Every ``stride``-th tile is executed: 20 is the reference's protocol (one tile per core, ref
    template_training.cpp:276-311); 1 runs all 20 tiles per core back to back, which on a GPU measures
    the per-tile cost at throughput instead of one launch latency (fit with ``--tiles-per-core 20``)."""
    xlen, ylen, zlen = shape
    if family == "fitcache":
        table, sequence = S.fitcache_stencils(variant), S.fitcache_sequence()
        outputs, constants, flush = ["t8"], ["i0"], False
    elif family == "fitddr":
        table, sequence = S.fitddr_stencils(variant), S.fitddr_sequence()
        outputs, flush = list(sequence), True
        constants = sorted({name for code in table.values() for name in _fields(code)} - set(sequence))
    else:
        raise ValueError(family)
    program = {
        "NAME": family, "CONSTANTS": constants, "OUTPUTS": outputs,
        "X": 2 * xlen, "Y": 2 * ylen, "Z": 5 * cores * zlen,
        "HX": 3, "HY": 3, "HZ": 3,
        "RUNS": 64, "VERIFY": False, "FLUSH": flush,
        "VARIANT": variant,
        "STENCILS": table, "SEQUENCE": sequence,
        "TILING": tiling_from_groups([sequence], [(2, 2, 5 * cores)]),
        "TRAINING": {"CORES": cores, "STRIDE": int(stride)},
    }
    return program


def _fields(code):
    from .analysis import stencil_accesses
    return [name for name, _ in stencil_accesses(code)]


def clone(program):
    """Deep copy: the analysis passes mutate the dict and are not idempotent."""
    return deepcopy(program)
