"""Kernel planning for the CUDA emitter.

Takes an analysed program (``analysis.analyse``) and decides, for every fused
group, everything ``template_tiling.cu`` needs to render one sm_100a kernel:

* the reference's tile geometry -- tile size ``T = ceil(S/N)`` and centred,
  truncating offset ``O = -(T*N - S)/2`` (ref template_tiling.cpp:228-233),
  tiles enumerated x-fastest (ref :240-242).  One tile is one CTA's unit of
  work; CTAs are persistent and take tiles ``blockIdx.x, +gridDim.x, ...``.
* a *sub-tile* ``B <= T``: the box whose haloed inputs (a ring of 2-4 TMA
  stages) and intermediates fit in shared memory.  A tile is walked as
  ``ceil(T/B)`` sub-tiles with overlapped (redundant) evaluation exactly like
  the tiles themselves, so results do not depend on ``B``.  With ``STREAM_K``
  the sub-tile is a column walked plane by plane through rings of k-planes.
* per input: the haloed box to stage (``LOOPS (+) OFFSETS`` of its consumers,
  what the reference's analyzer calls ``accesses``, ref stencil_analyzer.py:137-142),
  its shared-memory offset and strides, its tensor-map box.
* per *phase* (materialised stencils of equal box and dependency level): the
  evaluation box (tile grown by ``LOOPS``), where results go (shared-memory
  intermediate and/or global output) and straight-line code for a thread's
  ``1 x RJ x RK`` block of points in which every ``field(i+a,j+b,k+c)`` access
  is a memoised load, an inlined re-evaluation or a warp shuffle -- the
  expression text between accesses is untouched, so operand order is the
  reference's.

The plan also yields the launcher descriptor (``descriptor_text``): fields,
kernels, launch shapes, parameter-block layout, halo lists.
"""

import re

from . import analysis as A
from . import column

SMEM_LIMIT = 227 * 1024          # opt-in maximum per CTA on sm_100a
DEFAULTS = {
    "THREADS": 256,
    "STAGES": 2,
    "SMEM_BUDGET": 110 * 1024,   # two CTAs per SM
    "LOAD": "tma",               # "tma" | "ldg"
    "SUBTILE": None,             # (BX, BY, BZ) override
    "GROUPS": None,              # {"<kernel id>": {option: value}}: per-group overrides of the options above
    "FMAD": False,               # False -> -fmad=false: bit-exact with the strict oracle
    "FUSE_HALO": False,          # periodic copies written by the producing kernel
    "INLINE": "none",            # intermediates re-evaluated in registers: "none" | "all" | [names]
    "DIRECT": "none",            # centre-only inputs read straight from global: "none" | "auto" | [names]
    "BLOCK": (1, 1),             # (RJ, RK): points per thread along j and k
    "SHUFFLE": False,            # inlined intermediates at i-offsets come from neighbour lanes
    "BRANCH": False,             # expensive ?: arms become real (warp-coherent) branches
    "STORE_CS": False,           # outputs written with the evict-first (streaming) cache hint
    "L2_ONCE": False,            # centre-only inputs are staged with an L2 evict-first policy
    "POLL_WARP": False,          # only warp 0 polls the full-barrier of a stage, the other consumer warps wait in a
                                 # named barrier behind it (fewer issue slots burnt on polling); parity cases
                                 # "poll-warp*" in tests/test_parity_gpu.py, A/B in DESIGN.md
    "UNIFORM_WARP": False,       # warp index through __shfl_sync(.., 0): the compiler then knows that the role
                                 # split and the per-warp task loops are warp-uniform (uniform datapath, no
                                 # re-materialised descriptors in "divergent" code)
    "STORE_PTR": False,          # one global address per output and thread, the points of its block at constant offsets
                                 # (instead of one 64-bit index computation per point and output)
    "FAST_PATH": True,           # unclipped sub-tiles take a copy of the code with compile-time ranges (twice the code)
    "WAIT_SLEEP": 0,             # ns a starved consumer backs off between mbarrier polls
    "L2_PREFETCH": 0,            # work items whose boxes are prefetched into L2 ahead of the ring
    "STREAM_K": False,           # True: walk a tile column plane by plane through rings of k-planes in shared
                                 # memory; "regs": column kernel (column.py) -- a thread owns a column, values
                                 # wanted again at lower planes are carried in registers, no CTA barriers;
                                 # "auto": column kernel for groups whose stencils read each other across k
    "COLUMN": None,              # column kernel shape: {"WARPS": (WX, WY), "PLANES": planes per TMA box, "SLOTS": ring,
                                 # "UNROLL": ring slots per loop trip, "MERGE": walk whole columns of the kernel's own
                                 # shape instead of the variant's tiles, "CLUSTER": (CX, CY) tiles per thread-block
                                 # cluster whose producers stay within "SLACK" ring slots of each other (halo
                                 # sharing through L2), "WAVE_SYNC": see column.py}
    "PREFETCH": 2,               # STREAM_K: input planes in flight beyond the ones being read
}
FRONT_SLACK = 16                 # doubles before the first box (halo lanes of a warp read below it)
MAX_BOX = 256                    # TMA box limit per dimension
ALIGN_DOUBLES = 16               # 128-byte alignment of every staged box


def c_trunc_div(a, b):
    """C++ integer division (truncation toward zero) -- Python's // floors."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def tile_geometry(size, count):
    """(tile size, origin offset) of ``count`` tiles over ``size`` points (ref :228-233)."""
    t = A.ceil_div(size, count)
    return t, c_trunc_div(-(t * count - size), 2)


def options(program):
    opts = dict(DEFAULTS)
    opts.update(program.get("CUDA", {}))
    return opts


def _round_up(value, multiple):
    return (value + multiple - 1) // multiple * multiple


def _splits(extent):
    """Candidate sub-tile extents: even splits of ``extent`` into 1, 2, 3, ... parts."""
    seen, out = set(), []
    for parts in range(1, extent + 1):
        b = A.ceil_div(extent, parts)
        if b not in seen:
            seen.add(b)
            out.append(b)
    return out


def branchify(code, threshold=8):
    """Turn ``cond ? a : b`` with expensive arms into a real branch.

    C++ evaluates only the selected arm of ``?:`` anyway, so the values (and their rounding) are
    unchanged; what changes is that the compiler no longer if-converts the statement into evaluating
    *both* arms and selecting (advection: two 6-tap polynomials per direction, ref advection.py:19-54).
    With warp-coherent data -- wind direction does not flip between neighbouring columns -- the
    branch halves the fp64 work.  Cheap arms (the flux limiter of diffusion) stay selects.
    """
    out, counter = [], 0
    for statement in code.split(";"):
        while "?" in statement:
            q = statement.index("?")
            depth, start = 0, None
            for n in range(q - 1, -1, -1):               # start of the enclosing group
                c = statement[n]
                if c == ")":
                    depth += 1
                elif c == "(":
                    if depth == 0:
                        start = n + 1
                        break
                    depth -= 1
                elif c == "=" and depth == 0 and statement[n - 1] not in "<>!=" and statement[n + 1] != "=":
                    start = n + 1
                    break
            if start is None:
                break
            depth, colon, end = 0, None, len(statement)
            for n in range(q + 1, len(statement)):
                c = statement[n]
                if c == "(":
                    depth += 1
                elif c == ")":
                    if depth == 0:
                        end = n
                        break
                    depth -= 1
                elif c == ":" and depth == 0 and colon is None:
                    colon = n
                elif c == "?" and depth == 0:
                    colon = -1                       # nested conditional at the same level: leave it
                    break
            if colon is None or colon < 0:
                break
            cond, a, b = statement[start:q], statement[q + 1:colon], statement[colon + 1:end]
            if sum(a.count(op) + b.count(op) for op in "+-*/") < threshold:
                break
            name = "sel_%d" % counter
            counter += 1
            out.append("double %s; if (%s) { %s = %s; asm volatile(\"\"); } else { %s = %s; asm volatile(\"\"); }"
                       % (name, cond.strip(), name, a.strip(), name, b.strip()))
            statement = statement[:start] + " " + name + " " + statement[end:]
        out.append(statement)
    return ";".join(out)


class GroupPlan:
    """Everything needed to render and launch the kernel of one fused group.

    Stencils of the group fall into two classes.  *Materialised* stencils are evaluated over
    their LOOPS-extended box by all threads of the CTA and stored -- group outputs to global
    memory, intermediates that a later materialised stencil reads to shared memory.  *Inlined*
    intermediates (option ``INLINE``) are never stored: wherever a consumer reads them they are
    re-evaluated in registers from their own operands, memoised per thread.  Each thread owns a
    ``1 x RJ x RK`` block of points (option ``BLOCK``), so neighbouring evaluations share their
    loads and inlined values.  Materialised stencils with the same box that do not feed each
    other through shared memory are evaluated together in one *phase*; phases are separated by
    ``__syncthreads()`` where one reads what another wrote.
    """

    column = False

    def __init__(self, program, group, opts, kernel_name=None):
        self.program, self.group, self.opts = program, group, opts
        self.id = group["ID"] + 1          # kernel numbering follows the outer (schedule) id
        self.kernel_name = kernel_name or "absinthe_group%d" % self.id
        self.stencils = group["STENCILS"]
        self.names = [s["NAME"] for s in self.stencils]
        self.by_name = {s["NAME"]: s for s in self.stencils}
        dims = (program["X"], program["Y"], program["Z"])
        self.S = dims
        self.H = (program["HX"], program["HY"], program["HZ"])
        self.N = (group["NX"], group["NY"], group["NZ"])
        geo = [tile_geometry(s, n) for s, n in zip(dims, self.N)]
        self.T = tuple(g[0] for g in geo)
        self.O = tuple(g[1] for g in geo)
        self.tiles = self.N[0] * self.N[1] * self.N[2]
        self.reach = A.group_reach(group)
        self.inputs = list(group["INPUTS"])
        self.outputs = list(group["OUTPUTS"])
        self.row = dims[0] + 2 * self.H[0]
        self.plane = self.row * (dims[1] + 2 * self.H[1])
        self.block = (1,) + tuple(opts["BLOCK"])
        self._classify()
        self.load = opts["LOAD"]
        if self.load == "tma" and (self.row % 2 or not self.staged):
            self.load = "ldg"              # TMA needs 16-byte global strides
        self.stages = opts["STAGES"] if self.load == "tma" else 1
        self.threads = opts["THREADS"]
        self.stream = bool(opts["STREAM_K"]) and opts["STREAM_K"] != "regs" and self.load == "tma" and self.block[2] == 1
        if self.stream:
            # k-streaming: one work item is a (bx, by) column over the whole k-range of its tile.  Plane g
            # of every staged input travels together in ring slot g mod R; a phase whose results are read
            # up to d planes ahead runs d planes ahead of the output plane.
            self.prefetch = max(1, int(opts["PREFETCH"]))
            self.stages = self.prefetch + 1          # barrier slots: one full/empty pair per step in flight
            self.kmin = min(ph["BOX"][2][0] - ph["BOX"][2][1] for ph in self.phase_list)
            # planes of an input alive at output plane kk: [kk + lo, kk + hi], from the k-offsets its
            # consumers use and how far ahead of kk each consuming phase runs
            self.win = {}
            for ph in self.phase_list:
                ahead = ph["BOX"][2][1]
                for m in ph["NAMES"]:
                    for field, box in self.closure[m].items():
                        if field in self.staged:
                            lo, hi = ahead + box[2][0], ahead + box[2][1]
                            old = self.win.get(field, (lo, hi))
                            self.win[field] = (min(old[0], lo), max(old[1], hi))
        self.fuse_halo = bool(opts["FUSE_HALO"]) and all(s >= 2 * h for s, h in zip(dims, self.H))
        self.B = self._choose_subtile()
        self.nsub = tuple(A.ceil_div(t, b) for t, b in zip(self.T, self.B))
        self._layout()

    # -- classification ---------------------------------------------------------------------
    def _select(self, option, candidates):
        if option in (None, "none", False):
            return set()
        if option in ("all", True):
            return set(candidates)
        return set(option) & set(candidates)

    def _classify(self):
        temps = [n for n in self.names if n not in self.outputs]
        self.inlined = self._select(self.opts["INLINE"], temps)
        self.materialised = [n for n in self.names if n not in self.inlined]
        # fields a materialised stencil reads once its inlined operands are expanded
        self.closure = {}
        for name in self.names:
            reads = {}
            for field, box in self.by_name[name]["OFFSETS"].items():
                if field in self.inlined:
                    for inner, inner_box in self.closure[field].items():
                        grown = A.box_add(inner_box, box)
                        reads[inner] = A.box_widen(reads[inner], grown) if inner in reads else grown
                else:
                    reads[field] = A.box_widen(reads[field], box) if field in reads else box
            self.closure[name] = reads
        self.in_smem = [n for n in self.materialised
                        if any(n in self.closure[m] for m in self.materialised if m != n)]
        # a value that only feeds later stencils of its own phase at the same point stays in registers
        self.phase_list = self._phases()
        keep = []
        for n in self.in_smem:
            mine = next(ph for ph in self.phase_list if n in ph["NAMES"])
            if any(n in self.closure[m] for m in self.materialised if m != n and m not in mine["NAMES"]):
                keep.append(n)
        self.in_smem = keep
        centre_only = [n for n in self.inputs if self.reach[n] == A.ZERO_BOX]
        direct = "none" if self.opts["STREAM_K"] not in (False, None, "regs") else self.opts["DIRECT"]
        self.direct = self._select("all" if direct == "auto" else direct, centre_only)
        self.staged = [n for n in self.inputs if n not in self.direct]

    # -- shared-memory planning -----------------------------------------------------------
    def _box_dims(self, B, box, even_x=False, planes=False):
        d = [B[a] + box[a][1] - box[a][0] for a in range(3)]
        if planes and getattr(self, "stream", False):
            d[2] = 1
        if even_x and self.load == "tma":
            # TMA wants the box to start on a 16-byte boundary (an even fp64 column, measured:
            # odd start columns raise "illegal instruction") and its inner extent to be a
            # multiple of 16 bytes: load from the even column at or below the first one needed.
            d[0] = d[0] + 1 + (d[0] + 1) % 2
        return d

    def _slack(self, B):
        """Doubles a thread block may read past a box when its RJ x RK block overhangs the range."""
        worst = 0
        for name in self.staged + self.in_smem:
            d = self._box_dims(B, self.reach[name], even_x=name in self.staged, planes=True)
            worst = max(worst, (self.block[1] - 1) * d[0] + (self.block[2] - 1) * d[0] * d[1])
        if self.opts["SHUFFLE"]:
            worst += 48                     # lanes of the last warp segment past the end of a row
        return _round_up(worst, ALIGN_DOUBLES)

    def _temp_depth(self, name):
        """Planes of a streamed intermediate kept alive: its k-extent plus one (see template)."""
        return self.reach[name][2][1] - self.reach[name][2][0] + 2

    def _ring(self, name):
        """Planes of a streamed input resident at once: its k-window plus the prefetch distance."""
        lo, hi = self.win[name]
        return hi - lo + 1 + self.prefetch

    def _footprint(self, B):
        total = 0
        for name in self.staged:
            d = self._box_dims(B, self.reach[name], even_x=True, planes=True)
            depth = self._ring(name) if self.stream else self.stages
            total += depth * _round_up(d[0] * d[1] * d[2], ALIGN_DOUBLES)
        for name in self.in_smem:
            d = self._box_dims(B, self.reach[name], planes=True)
            depth = self._temp_depth(name) if self.stream else 1
            total += depth * _round_up(d[0] * d[1] * d[2], ALIGN_DOUBLES)
        return (total + self._slack(B) + FRONT_SLACK) * 8 + 128 + 16 * self.stages

    def _cost(self, B):
        """Work per useful point of a sub-tile ``B``: staged cells (short rows cost extra: every
        row pays about a sector and a half of slack), evaluated cells weighted by their operand
        count, and a fixed price per work item (barriers, pipeline bubble)."""
        row_slack, item_overhead = 12, 3000
        moved = item_overhead
        for name in self.staged:
            d = self._box_dims(B, self.reach[name], even_x=True, planes=True)
            moved += (d[0] + row_slack) * d[1] * d[2]
        for name in self.materialised:
            d = self._box_dims(B, self.reach[name], planes=True)
            moved += d[0] * d[1] * d[2] * 0.25 * len(self.closure[name])
        return moved / float(B[0] * B[1] * (1 if self.stream else B[2]))

    def _choose_subtile(self):
        if self.opts["SUBTILE"]:
            B = tuple(min(b, t) for b, t in zip(self.opts["SUBTILE"], self.T))
            if self.stream:
                B = (B[0], B[1], self.T[2])
            assert self._footprint(B) <= SMEM_LIMIT, "SUBTILE does not fit in shared memory"
            return B
        widest = [0, 0, 0]
        for name in self.inputs + self.names:
            for a in range(3):
                widest[a] = max(widest[a], self.reach[name][a][1] - self.reach[name][a][0])
        budget = min(self.opts["SMEM_BUDGET"], SMEM_LIMIT)
        best = None
        for bx in _splits(self.T[0]):
            if bx + widest[0] + 2 > MAX_BOX:
                continue
            for by in _splits(self.T[1]):
                if by + widest[1] > MAX_BOX:
                    continue
                for bz in ([self.T[2]] if self.stream else _splits(self.T[2])):
                    if not self.stream and bz + widest[2] > MAX_BOX:
                        continue
                    B = (bx, by, bz)
                    # a thread's RJ x RK block should not overhang the sub-tile
                    if any(B[a] % self.block[a] and B[a] != self.T[a] and self.T[a] >= self.block[a]
                           for a in (1, 2)):
                        continue
                    if self._footprint(B) > budget:
                        continue
                    key = (round(self._cost(B), 3), -bx * by * bz)
                    if best is None or key < best[0]:
                        best = (key, B)
        if best is None:
            B = (min(self.T[0], 32), 1, 1)
            assert self._footprint(B) <= SMEM_LIMIT, "group does not fit in shared memory"
            return B
        return best[1]

    def _layout(self):
        B = self.B
        cursor = 0
        self.front = FRONT_SLACK if self.opts["SHUFFLE"] else 0
        self.input_plan = []
        for n, name in enumerate(self.inputs):
            if name in self.direct:
                continue
            box = self.reach[name]
            d = self._box_dims(B, box, even_x=True, planes=True)
            lo = [box[a][0] for a in range(3)]
            self.input_plan.append({
                "NAME": name, "INDEX": n, "MAP": len(self.input_plan), "LO": lo, "DIM": d, "OFFSET": cursor,
                "ROW": d[0], "PLANE": d[0] * d[1], "CELLS": d[0] * d[1] * d[2],
                # pointer bias so that relative (i, j, k) index the box directly
                "BIAS": cursor - (lo[0] + lo[1] * d[0] + (0 if self.stream else lo[2] * d[0] * d[1])),
                "ONCE": bool(self.opts["L2_ONCE"]) and box == A.ZERO_BOX,
            })
            slot = _round_up(d[0] * d[1] * d[2], ALIGN_DOUBLES)
            if self.stream:
                ring = self._ring(name)
                self.input_plan[-1].update(RING=ring, SLOT=slot, WK=self.win[name][1] - self.win[name][0],
                                           LOKX=self.win[name][0])
                cursor += ring * slot
            else:
                cursor += slot
        self.stage_doubles = cursor
        self.stage_bytes = sum(p["CELLS"] for p in self.input_plan) * 8
        if self.stream:
            self.first_bytes = sum(p["CELLS"] * (p["WK"] + 1) for p in self.input_plan) * 8
        else:
            cursor = self.stages * self.stage_doubles
        self.temp_plan = {}
        for name in self.in_smem:
            box = self.reach[name]
            d = self._box_dims(B, box, planes=True)
            lo = [box[a][0] for a in range(3)]
            depth = self._temp_depth(name) if self.stream else 1
            slot = _round_up(d[0] * d[1] * d[2], ALIGN_DOUBLES)
            self.temp_plan[name] = {
                "NAME": name, "LO": lo, "DIM": d, "OFFSET": cursor, "ROW": d[0], "PLANE": d[0] * d[1],
                "BIAS": cursor - (lo[0] + lo[1] * d[0] + (0 if self.stream else lo[2] * d[0] * d[1])),
                "DEPTH": depth, "SLOT": slot,
            }
            cursor += depth * slot
        cursor += self._slack(B) + self.front
        self.barrier_offset = cursor * 8
        self.smem_bytes = cursor * 8 + 16 * max(self.stages, 1) + 48

        self.launch_threads = self.threads + (32 if self.load == "tma" else 0)
        assert self.smem_bytes <= SMEM_LIMIT, "shared-memory plan exceeds the sm_100a limit"
        # parameter block: tensor maps (128 B each, 64-byte aligned), pointers, tile control
        off = 0
        self.args = []
        if self.load == "tma":
            for p in self.input_plan:
                self.args.append(("tmap", off, p["NAME"], p["DIM"]))
                off += 128
        for name in self.inputs:
            self.args.append(("ptr", off, name, None))
            off += 8
        if not self.inputs:
            off += 8                        # the kernel declares at least one input slot
        for name in self.outputs:
            self.args.append(("ptr", off, name, None))
            off += 8
        self.args.append(("tilectl", off, None, None))
        off += 8
        if self.fuse_halo:
            self.args.append(("remote", off, None, None))    # neighbours' exchange windows (multi-GPU fused PUT)
            off += 16
        self.params_bytes = _round_up(off, 64)

    # -- phases -----------------------------------------------------------------------------
    def _phases(self):
        """Group the materialised stencils into barrier-separated phases.

        A stencil may share a phase with a producer it reads only at its own point (the value
        is still in the thread's registers) when both cover the same box; every other
        shared-memory dependence puts it one level later.  Stencils of equal level and box form
        one phase, so e.g. the four laplacians of a fused diffusion group are one sweep.
        """
        level = {}
        for name in self.materialised:
            box, lv = self.reach[name], 0
            direct_zero = {f for f, b in self.by_name[name]["OFFSETS"].items() if b == A.ZERO_BOX}
            for f in self.closure[name]:
                if f not in level:
                    continue
                free = (f in direct_zero and self._only_direct(name, f) and self.reach[f] == box)
                lv = max(lv, level[f] + (0 if free else 1))
            level[name] = lv
        phases = {}
        for name in self.materialised:
            phases.setdefault((level[name], self.reach[name]), []).append(name)
        order = sorted(phases, key=lambda key: (key[0], self.names.index(phases[key][0])))
        return [{"BOX": key[1], "NAMES": phases[key]} for key in order]

    def _only_direct(self, name, field):
        """True when ``name`` reaches ``field`` only by its own zero-offset access (no inlined path)."""
        for other in self.by_name[name]["OFFSETS"]:
            if other in self.inlined and field in self.closure[other]:
                return False
        return True

    @staticmethod
    def _tag(off):
        return "_".join(("m%d" % -o) if o < 0 else str(o) for o in off)

    def _phase_body(self, phase):
        """Straight-line code for one thread: its RJ x RK block of every stencil of the phase.

        Returns (lines, stores, lanes).  With SHUFFLE an inlined intermediate wanted at an
        i-offset is taken from the neighbour lane that evaluated it at its own column;
        ``lanes = (first, last)`` are the lanes of a warp whose results are then complete.
        """
        staged = {p["NAME"]: p for p in self.input_plan}
        shuffle = bool(self.opts["SHUFFLE"])
        lines, memo, lanes = [], {}, {}
        same_phase = set(phase["NAMES"])
        planes = phase.setdefault("PLANES", set())     # (kind, field, k-offset) plane pointers needed

        def const_index(off, row, plane):
            return off[0] + off[1] * row + (0 if self.stream else off[2] * plane)

        def base(kind, field, off):
            """Name of the array a load goes through: one pointer per k-plane when streaming."""
            if not self.stream:
                return "%s_%s" % (kind, field)
            planes.add((kind, field, off[2]))
            return "%s_%s_k%s" % (kind, field, ("m%d" % -off[2]) if off[2] < 0 else str(off[2]))

        def value(field, off):
            """C expression of ``field`` at block-relative offset ``off``."""
            key = (field, off)
            if key in memo:
                return memo[key]
            tag = "%s_%s" % (field, self._tag(off))
            if field in staged:
                p = staged[field]
                var = "l_" + tag
                lines.append("const double %s = %s[b_%s + (%d)];" % (
                    var, base("in", field, off), field, const_index(off, p["ROW"], p["PLANE"])))
                lanes[var] = (0, 31)
            elif field in self.direct:
                var = "l_" + tag
                assert off[0] == 0, "direct inputs are centre-only"
                lines.append("const double %s = ok_%d_%d ? __ldg(gin_%s + bg + (%dL)) : 0.0;" % (
                    var, off[1], off[2], field, off[1] * self.row + off[2] * self.plane))
                lanes[var] = (0, 31)
            elif field in self.temp_plan and field not in same_phase:
                p = self.temp_plan[field]
                var = "l_" + tag
                lines.append("const double %s = %s[b_%s + (%d)];" % (
                    var, base("tmp", field, off), field, const_index(off, p["ROW"], p["PLANE"])))
                lanes[var] = (0, 31)
            elif field in same_phase:
                raise AssertionError("phase split missed a dependence on " + field)
            else:
                assert field in self.inlined, "stencil reads unknown field " + field
                if shuffle and off[0] != 0:
                    own = value(field, (0, off[1], off[2]))
                    var = "s_" + tag
                    fn = "__shfl_down_sync" if off[0] > 0 else "__shfl_up_sync"
                    lines.append("const double %s = %s(0xffffffffu, %s, %d);" % (var, fn, own, abs(off[0])))
                    lo, hi = lanes[own]
                    lanes[var] = (max(0, lo - off[0]), min(31, hi - off[0]))
                else:
                    var = evaluate(field, off)
            memo[key] = var
            return var

        def evaluate(name, off):
            """Emit ``name`` evaluated at ``off``; returns the variable holding its result."""
            code = self.by_name[name]["LAMBDA"]
            lo, hi = 0, 31
            for field, d in A.stencil_accesses(code):       # operands first (emits their code)
                operand = value(field, tuple(o + e for o, e in zip(off, d)))
                lo, hi = max(lo, lanes[operand][0]), min(hi, lanes[operand][1])

            def sub(match):
                d = tuple(A._offset(match.group(n)) for n in (2, 3, 4))
                return memo[(match.group(1), tuple(o + e for o, e in zip(off, d)))]
            var = "v_%s_%s" % (name, self._tag(off))
            body = A._ACCESS.sub(sub, code)
            if self.opts["BRANCH"]:
                body = branchify(body)
            lines.append("double %s; { %s %s = res; }" % (var, body, var))
            lanes[var] = (lo, hi)
            return var

        stores = []
        first, last = 0, 31
        for rk in range(self.block[2]):
            for rj in range(self.block[1]):
                point = (0, rj, rk)
                for name in phase["NAMES"]:
                    var = evaluate(name, point)
                    memo[(name, point)] = var
                    stores.append((name, rj, rk, var))
                    first, last = max(first, lanes[var][0]), min(last, lanes[var][1])
        assert first <= last, "shuffle chain longer than a warp"
        # lanes left of the first owning one read below the start of a staged row: covered by FRONT_SLACK
        assert first + max(self.H) + 1 <= FRONT_SLACK, "shuffle chain reaches below the front slack"
        return lines, stores, (first, last)

    def passes(self):
        out = []
        pending = set()
        for phase in self.phase_list:
            box = phase["BOX"]
            lo = [box[a][0] for a in range(3)]
            hi = [box[a][1] for a in range(3)]
            dim = self._box_dims(self.B, box, planes=True)
            blocks = [dim[0], A.ceil_div(dim[1], self.block[1]), A.ceil_div(dim[2], self.block[2])]
            reads = set()
            for name in phase["NAMES"]:
                reads.update(f for f in self.closure[name] if f in self.temp_plan and f not in phase["NAMES"])
            sync_before = bool(reads & pending)
            if sync_before:
                pending.clear()
            pending.update(n for n in phase["NAMES"] if n in self.temp_plan)
            lines, stores, lanes = self._phase_body(phase)
            useful = lanes[1] - lanes[0] + 1
            nseg = A.ceil_div(dim[0], useful)
            pointers = []
            if self.stream:
                wanted = set(phase["PLANES"]) | {("tmp", n, 0) for n in phase["NAMES"] if n in self.temp_plan}
                staged = {p["NAME"]: p for p in self.input_plan}
                for kind, field, c in sorted(wanted):
                    plan = staged[field] if kind == "in" else self.temp_plan[field]
                    pointers.append({"KIND": kind, "FIELD": field, "OFF": c, "PLAN": plan,
                                     "NAME": "%s_%s_k%s" % (kind, field, ("m%d" % -c) if c < 0 else str(c))})
            out.append({
                "D": hi[2], "KLO": lo[2], "POINTERS": pointers,
                "SHUFFLE": bool(self.opts["SHUFFLE"]), "LANES": lanes, "USEFUL": useful, "NSEG": nseg,
                "TASKS": nseg * blocks[1] * blocks[2],
                "NAMES": phase["NAMES"], "LO": lo, "HI": hi, "DIM": dim, "BLOCKS": blocks,
                "CELLS": blocks[0] * blocks[1] * blocks[2], "SYNC_BEFORE": sync_before,
                "BODY": lines,
                "STORES": [{"NAME": n, "RJ": rj, "RK": rk, "VAR": var,
                            "TEMP": self.temp_plan.get(n), "TO_GLOBAL": n in self.outputs} for n, rj, rk, var in stores],
                "OUTPUTS": [{"NAME": n, "INDEX": self.outputs.index(n)} for n in phase["NAMES"] if n in self.outputs],
                "READS_STAGED": sorted({f for n in phase["NAMES"] for f in self.closure[n] if f in self.staged}),
                "READS_DIRECT": sorted({f for n in phase["NAMES"] for f in self.closure[n] if f in self.direct}),
                "READS_TEMPS": sorted(reads),
                "POINTS": [(rj, rk) for rk in range(self.block[2]) for rj in range(self.block[1])],
            })
        return out

    def context(self):
        return {
            "ID": self.id, "KERNEL": self.kernel_name,
            "THREADS": self.threads, "STAGES": self.stages, "TMA": self.load == "tma",
            "T": self.T, "O": self.O, "N": self.N, "B": self.B, "NSUB": self.nsub,
            "NSUB_TOTAL": self.nsub[0] * self.nsub[1] * self.nsub[2],
            "TILES": self.tiles, "INPUTS": self.input_plan, "NIN": len(self.inputs),
            "NMAPS": len(self.input_plan),
            "DIRECT": [{"NAME": n, "INDEX": self.inputs.index(n)} for n in self.inputs if n in self.direct],
            "OUTPUTS": self.outputs, "NOUT": len(self.outputs),
            "TEMPS": list(self.temp_plan.values()), "PASSES": self.passes(),
            "STAGE_DOUBLES": self.stage_doubles, "STAGE_BYTES": self.stage_bytes,
            "BARRIER_OFFSET": self.barrier_offset, "SMEM_BYTES": self.smem_bytes,
            "PARAMS_BYTES": self.params_bytes, "NAMES": self.names, "INLINED": sorted(self.inlined),
            "BLOCK": self.block, "FUSE_HALO": self.fuse_halo, "FRONT": self.front,
            "WARPS": self.threads // 32, "LAUNCH_THREADS": self.launch_threads,
            "FAST_PATH": bool(self.opts["FAST_PATH"]), "STORE_PTR": bool(self.opts["STORE_PTR"]), "STREAM_STORES": bool(self.opts["STORE_CS"]), "UNIFORM_WARP": bool(self.opts["UNIFORM_WARP"]), "POLL_WARP": bool(self.opts["POLL_WARP"]), "L2_PREFETCH": int(self.opts["L2_PREFETCH"]), "STREAM": self.stream, "KMIN": getattr(self, "kmin", 0), "PREFETCH": getattr(self, "prefetch", 0),
            "FIRST_BYTES": getattr(self, "first_bytes", 0), "NB": self.stages,
        }


class Plan:
    """Kernel plans of all fused groups of one analysed program."""

    def __init__(self, program):
        A.analyse(program)
        self.program = program
        self.opts = options(program)
        self.groups = []
        for outer in program["TILING"]["GROUPS"]:
            if not outer["LOOPS"]:
                continue                      # the dummy input-halo group
            for inner in outer["GROUPS"]:
                opts = self._group_options(inner["ID"] + 1)
                # kernels carry the sequence's name so that launch lists attribute them (absinthe_<name>_g<id>)
                symbol = "absinthe_%s_g%d" % (re.sub(r"\W", "_", str(program.get("NAME", "variant"))), inner["ID"] + 1)
                wanted = opts["STREAM_K"] == "regs" or (opts["STREAM_K"] == "auto" and column.k_dependent(inner))
                if opts["STREAM_K"] == "auto" and not wanted:
                    opts = dict(opts, STREAM_K=False)
                if wanted and column.applicable(program, inner, opts):
                    self.groups.append(column.ColumnPlan(program, inner, opts, symbol))
                else:
                    self.groups.append(GroupPlan(program, inner, opts, symbol))
        tiling = program["TILING"]
        self.inputs = sorted(tiling["INPUTS"])
        self.outputs = sorted(tiling["OUTPUTS"])
        produced = [n for g in self.groups for n in g.outputs]
        self.temps = sorted(set(produced) - set(self.outputs))

    def _group_options(self, kernel_id):
        """The options of one group kernel: the program-wide ones, then ``GROUPS[<kernel id>]`` on top
        (launch shape, staging and sub-tile choices differ between the groups of one variant)."""
        per_group = self.opts.get("GROUPS") or {}
        override = per_group.get(str(kernel_id), per_group.get(kernel_id, {}))
        unknown = set(override) - set(DEFAULTS) | set(override) & {"GROUPS", "FMAD", "WAIT_SLEEP"}
        assert not unknown, "not a per-group option: %s" % sorted(unknown)
        opts = dict(self.opts)
        opts.update(override)
        return opts

    def context(self):
        p = self.program
        ctx = {k: p[k] for k in ("X", "Y", "Z", "HX", "HY", "HZ")}
        ctx.update(VARIANT=p.get("VARIANT", p.get("NAME", "")), GROUPS=[g.context() for g in self.groups],
                   ANY_TMA=any(g.load == "tma" for g in self.groups), FMAD=self.opts["FMAD"],
                   WAIT_SLEEP=int(self.opts["WAIT_SLEEP"]))
        return ctx

    def nvcc_flags(self):
        flags = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a"]
        if not self.opts["FMAD"]:
            flags.append("-fmad=false")
        return flags

    def descriptor_text(self):
        p = self.program
        training = p.get("TRAINING", {})
        lines = ["absinthe-variant 1", "name %s" % (p.get("VARIANT") or p.get("NAME") or "variant"),
                 "# runs %d" % p.get("RUNS", 64), "# flush %d" % int(bool(p.get("FLUSH", True))),
                 "# training %d" % int(bool(training)),
                 # node grid of the variant description (ref template_tiling.cpp:189-203): ranks along j; the
                 # domain line below is one rank's slab, "global" the decomposed extent
                 "# nodes %d %d %d" % tuple(p["TILING"].get(k, 1) for k in ("NX", "NY", "NZ")),
                 "# global %d %d %d" % tuple(p.get("GLOBAL", (p["X"], p["Y"], p["Z"])))]
        lines += ["# verify 1"] if p.get("VERIFY") else []
        lines.append("domain %d %d %d %d %d %d" % tuple(p[k] for k in ("X", "Y", "Z", "HX", "HY", "HZ")))
        for name in self.inputs:
            lines.append("field %s in" % name)
        for name in self.outputs:
            lines.append("field %s out" % name)
        for name in self.temps:
            lines.append("field %s tmp" % name)
        for g in self.groups:
            lines.append("group %d kernel %s threads %d smem %d tiles %d params %d "
                         "training_step %d fused_halo %d%s%s" % (
                             g.id, g.kernel_name, g.launch_threads, g.smem_bytes, g.tiles, g.params_bytes,
                             training.get("STRIDE", 20), int(g.fuse_halo),
                             " scratch %d" % g.scratch_bytes if getattr(g, "scratch_bytes", 0) else "",
                             " cluster %d" % g.cluster_size if getattr(g, "cluster_size", 1) > 1 else ""))
            for kind, off, name, dim in g.args:
                if kind == "tmap":
                    lines.append("arg tmap %d %s %d %d %d" % (off, name, dim[0], dim[1], dim[2]))
                elif kind == "ptr":
                    lines.append("arg ptr %d %s" % (off, name))
                elif kind == "scratch":
                    lines.append("arg scratch %d" % off)
                elif kind == "remote":
                    lines.append("arg remote %d" % off)
                else:
                    lines.append("arg tilectl %d" % off)
            for name in g.outputs:
                lines.append("halo %s" % name)
            lines.append("endgroup")
        lines.append("end")
        return "\n".join(lines) + "\n"
