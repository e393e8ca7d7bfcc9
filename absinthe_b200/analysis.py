"""Stencil-program analysis: access boxes, group dataflow, overlapped-tiling extents.

This is the B200 backend's own analyzer.  It produces the *same variant
description* (the program dict after analysis) that the reference's
``stencil_analyzer.py`` produces, because that dict is the kernel
specification the emitter consumes:

* per stencil  ``{"NAME", "OFFSETS", "LAMBDA"}``        (ref stencil_analyzer.py:112-125)
* per group    ``INPUTS / OUTPUTS / TEMPS``             (ref stencil_analyzer.py:194-237)
* per group    ``LOOPS`` (redundant-compute extents) and ``HALOS``
                                                        (ref stencil_analyzer.py:128-156)
* a leading dummy group carrying the program-input halos, group ``ID``s
                                                        (ref stencil_analyzer.py:159-191)

The public entry points keep the reference's names (``verify_program``,
``compute_dataflow``, ``compute_boundaries``, ``analyze_stencil``,
``count_fetches``) and mutate the program dict in place like the reference
does.  One deliberate difference: set-valued results (INPUTS/OUTPUTS/TEMPS)
are emitted in *sorted* order, so the generated code and the input
initialisation order do not depend on ``PYTHONHASHSEED`` (SURVEY.md §7,
"Input reproducibility").  ``tests/test_analysis.py`` checks the dicts
against fixtures dumped from the reference analyzer (as sets).
"""

import re

ZERO_BOX = ((0, 0), (0, 0), (0, 0))

# name(i<off>, j<off>, k<off>) with an optional single-digit signed offset
# per index -- the access grammar of the reference's stencil strings
# (ref stencil_analyzer.py:71-85).
_ACCESS = re.compile(
    r"(\w+)\(\s*i+(\s*[+-]+\s*\d)?\s*,\s*j+(\s*[+-]+\s*\d)?\s*,\s*k+(\s*[+-]+\s*\d)?\s*\)")


def _offset(text):
    return int(text.replace(" ", "")) if text else 0


def stencil_accesses(code):
    """All ``(field, (di, dj, dk))`` accesses of a stencil string, in text order."""
    return [(m.group(1), (_offset(m.group(2)), _offset(m.group(3)), _offset(m.group(4))))
            for m in _ACCESS.finditer(code)]


def parse_stencil(code):
    """Map ``field -> sorted list of offsets`` (duplicates kept, like the reference)."""
    table = {}
    for name, off in sorted(stencil_accesses(code)):
        table.setdefault(name, []).append(off)
    return table


def analyze_stencil(code):
    """Map ``field -> ((imin, imax), (jmin, jmax), (kmin, kmax))`` of a stencil string."""
    boxes = {}
    for name, offs in parse_stencil(code).items():
        boxes[name] = tuple((min(o[d] for o in offs), max(o[d] for o in offs)) for d in range(3))
    return boxes


def count_fetches(code):
    """Distinct operand fetches of a stencil plus one for its write (ref :101-109)."""
    return 1 + sum(len(set(offs)) for offs in parse_stencil(code).values())


# -- box algebra -------------------------------------------------------------------------
# A box is ((ilo, ihi), (jlo, jhi), (klo, khi)).  ``_pick`` reproduces the reference's
# extension rule (min when the pair sums to <= 0, max otherwise) so that even the
# unusual all-positive-lower-bound case gives identical LOOPS.

def _pick(a, b, widen):
    nonpos = (a + b) <= 0
    if widen:
        return min(a, b) if nonpos else max(a, b)
    return max(a, b) if nonpos else min(a, b)


def box_add(p, q):
    return tuple((p[d][0] + q[d][0], p[d][1] + q[d][1]) for d in range(3))


def box_widen(p, q):
    return tuple((_pick(p[d][0], q[d][0], True), _pick(p[d][1], q[d][1], True)) for d in range(3))


def box_narrow(p, q):
    return tuple((_pick(p[d][0], q[d][0], False), _pick(p[d][1], q[d][1], False)) for d in range(3))


def _halo_record(remote, local):
    inner = box_narrow(remote, local)
    return {"OX": remote[0], "IX": inner[0], "OY": remote[1], "IY": inner[1],
            "OZ": remote[2], "IZ": inner[2]}


def _halo_needed(rec):
    for o, i in (("OX", "IX"), ("OY", "IY"), ("OZ", "IZ")):
        if rec[o][0] - rec[i][0] < 0 or rec[o][1] - rec[i][1] > 0:
            return True
    return False


def _fits_halo(boxes, program):
    limit = (program["HX"], program["HY"], program["HZ"])
    return all(max(abs(b[d][0]), abs(b[d][1])) <= limit[d] for b in boxes.values() for d in range(3))


# -- dataflow ----------------------------------------------------------------------------

def _classify(group, steps, needed):
    """steps: [(reads, writes)] in execution order; needed: fields required downstream."""
    reads_all, outs, temps = set(), set(), set()
    for reads, writes in steps:
        reads_all.update(reads)
        for w in writes:
            (outs if w in needed else temps).add(w)
    ins = reads_all - outs - temps
    group["OUTPUTS"] = sorted(outs)
    group["INPUTS"] = sorted(ins)
    group["TEMPS"] = sorted(temps)
    return needed | ins


def _annotate_stencils(program):
    for outer in program["TILING"]["GROUPS"]:
        for inner in outer["GROUPS"]:
            inner["STENCILS"] = [
                {"NAME": s, "OFFSETS": analyze_stencil(program["STENCILS"][s]),
                 "LAMBDA": program["STENCILS"][s]} if not isinstance(s, dict) else s
                for s in inner["STENCILS"]]


def compute_dataflow(program):
    """Fill INPUTS/OUTPUTS/TEMPS at the inner-group, outer-group and program level."""
    _annotate_stencils(program)
    needed_outer = set(program["OUTPUTS"])
    for outer in reversed(program["TILING"]["GROUPS"]):
        needed_inner = set(needed_outer)
        for inner in reversed(outer["GROUPS"]):
            steps = [(list(s["OFFSETS"]), [s["NAME"]]) for s in inner["STENCILS"]]
            needed_inner = _classify(inner, steps, needed_inner)
        needed_outer = _classify(outer, [(g["INPUTS"], g["OUTPUTS"]) for g in outer["GROUPS"]],
                                 needed_outer)
    steps = [(g["INPUTS"], g["OUTPUTS"]) for g in program["TILING"]["GROUPS"]]
    _classify(program["TILING"], steps, set(program["OUTPUTS"]))


# -- boundaries --------------------------------------------------------------------------

def _extents(program, group, stencils, downstream):
    """Backward sweep over ``stencils`` accumulating how far past the tile each field is read."""
    reach = {name: ZERO_BOX for name in group["OUTPUTS"]}
    for name, offsets in reversed(stencils):
        for field, box in offsets.items():
            grown = box_add(box, reach[name])
            reach[field] = box_widen(reach[field], grown) if field in reach else grown
    assert _fits_halo(reach, program), "accesses too wide"
    group["LOOPS"] = {name: reach[name] for name, _ in stencils}
    group["HALOS"] = {}
    for name in group["OUTPUTS"]:
        rec = _halo_record(downstream[name], reach[name])
        if _halo_needed(rec):
            group["HALOS"][name] = rec
    for name in group["INPUTS"]:
        downstream[name] = box_widen(downstream[name], reach[name]) if name in downstream else reach[name]
    return downstream


def compute_boundaries(program):
    """Fill LOOPS/HALOS, prepend the dummy input-halo group and number the groups."""
    outer_dep = {name: ZERO_BOX for name in program["OUTPUTS"]}
    for outer in reversed(program["TILING"]["GROUPS"]):
        inner_dep = {name: ZERO_BOX for name in outer["OUTPUTS"]}
        flat = []
        for inner in reversed(outer["GROUPS"]):
            mine = [(s["NAME"], s["OFFSETS"]) for s in inner["STENCILS"]]
            inner_dep = _extents(program, inner, mine, inner_dep)
            flat = mine + flat
        outer_dep = _extents(program, outer, flat, outer_dep)
    halos = {}
    for name in program["TILING"]["INPUTS"]:
        rec = _halo_record(outer_dep[name], ZERO_BOX)
        if _halo_needed(rec):
            halos[name] = rec
    head = {"TEMPS": [], "INPUTS": [], "OUTPUTS": [], "GROUPS": [], "LOOPS": {}, "HALOS": halos}
    program["TILING"]["GROUPS"] = [head] + program["TILING"]["GROUPS"]
    inner_id = 0
    for outer_id, outer in enumerate(program["TILING"]["GROUPS"]):
        outer["ID"] = outer_id
        for inner in outer["GROUPS"]:
            inner["ID"] = inner_id
            inner_id += 1


def group_reach(group):
    """How far beyond its tile a fused group touches every field it reads or writes.

    Same backward sweep as ``_extents``; returns ``field -> box`` for the group's inputs
    (the haloed tile a kernel must stage) and for its stencils (equal to ``LOOPS``).
    """
    reach = {name: ZERO_BOX for name in group["OUTPUTS"]}
    for stencil in reversed(group["STENCILS"]):
        for field, box in stencil["OFFSETS"].items():
            grown = box_add(box, reach[stencil["NAME"]])
            reach[field] = box_widen(reach[field], grown) if field in reach else grown
    return reach


def ceil_div(a, b):
    return (a + b - 1) // b


def verify_program(program):
    """Reject decompositions whose subdomains or tiles would be empty (ref :240-259)."""
    tiling = program["TILING"]
    sub = {}
    for dim in "XYZ":
        sub[dim] = ceil_div(program[dim], tiling["N" + dim])
        assert sub[dim] > 0, dim.lower() + " size not large enough for domain decomposition"
    for outer in tiling["GROUPS"]:
        for inner in outer["GROUPS"]:
            for dim in "XYZ":
                assert ceil_div(sub[dim], inner["N" + dim]) > 0, \
                    dim.lower() + " size not large enough for cache tiling"


def is_analysed(program):
    groups = program["TILING"]["GROUPS"]
    return bool(groups) and "ID" in groups[0]


def analyse(program):
    """verify + dataflow + boundaries, once (the reference's passes are not idempotent)."""
    if not is_analysed(program):
        verify_program(program)
        compute_dataflow(program)
        compute_boundaries(program)
    return program


def algorithmic_bytes_per_point(program):
    """8 B x sum over fusion groups of (|INPUTS| + |OUTPUTS|)  (SURVEY.md section 8d)."""
    total = 0
    for outer in program["TILING"]["GROUPS"]:
        for inner in outer["GROUPS"]:
            total += len(inner["INPUTS"]) + len(inner["OUTPUTS"])
    return 8 * total
