"""Column kernels: register-carried streaming along k, one warp per column strip, no CTA barriers.

The fused tile loop of the reference (ref template_tiling.cpp:301-336) walks a tile stencil by
stencil; a group with k-dependences (fast waves: ``ppgc`` reads ``ppgk(k+1)``, ``udc`` reads
``uout(k-1)``, ``div`` reads ``udc(k+1)`` -- ref fastwaves.py:16-48) then needs every intermediate
on a k-extended box.  On the GPU that is what made large k-dependent groups slow: re-evaluating the
chain per output plane triples the divisions, and keeping plane rings of intermediates in shared
memory costs a CTA barrier per phase and plane.  A *column* kernel removes both:

* a work item is a ``bx x by`` column of a tile over the tile's k-range; a thread owns one column
  ``i`` and ``RJ`` consecutive rows, and walks k.  A warp is a strip of 32 consecutive columns;
* every stencil ``S`` runs ``lead[S]`` planes ahead of the output plane ``kk`` (software pipelining
  along k: ``lead[S] >= lead[C] + c`` for every consumer ``C`` reading ``S(k+c)``), so at step ``kk``
  it is evaluated exactly once, at plane ``kk + lead[S]``;
* values wanted again at lower planes in later steps are *carried in registers* from step to step
  (``c_<key>_q<q>`` holds the value at plane ``kk + q``) -- nothing is evaluated twice along k;
* an intermediate wanted at an i-offset comes from the neighbour lane that owns that column
  (``__shfl_up/down_sync``); at a j-offset it comes from the thread's own rows: a thread evaluates
  the rows ``rows[S]`` its outputs need (``RJ`` plus the j-reach of the consumers);
* staged inputs are read from the newest plane of their box only (the ring holds one plane window;
  lower planes are carried like intermediates), so a ring slot can be released as soon as its last
  plane has been consumed and almost the whole ring is in flight;
* the TMA producer warp delivers, per ring slot, one ``DX x DY x KC`` box per input (``KC`` planes per
  request); consumer warps meet the producer through ``full[]/empty[]`` mbarriers once per slot and
  never meet each other;
* optionally (``COLUMN.CLUSTER``) the kernel runs in thread-block clusters: the CTAs of a cluster take
  neighbouring tiles and their producers pace each other through remote ``mbarrier`` arrives, so the
  halo cells two columns share are fetched from HBM once (``_cluster`` below).

The plan is built as a small IR (``ops`` + ``carries``) from which the CUDA text is rendered;
``simulate`` replays the IR on symbolic tags (which stencil / column / row / plane a register holds),
which is how the schedule is verified on a machine without a GPU (tests/test_column.py).
"""

from . import analysis as A

SMEM_LIMIT = 227 * 1024
FRONT_SLACK = 16
TAIL_SLACK = 64
ALIGN_DOUBLES = 16
MAX_BOX = 256
SYNC_LINES = 4                   # counter lines per CTA of the wave barrier (work items per CTA it can cover: 4 x grid)


def _round_up(value, multiple):
    return (value + multiple - 1) // multiple * multiple


def _tag(value):
    return ("m%d" % -value) if value < 0 else str(value)


def applicable(program, group, opts):
    """Column kernels need TMA-able rows (16-byte global strides) and one point per thread along k."""
    row = program["X"] + 2 * program["HX"]
    return opts.get("LOAD", "tma") == "tma" and row % 2 == 0 and bool(group["INPUTS"])


def k_dependent(group):
    """True when a stencil of the group reads another stencil of the group at a k-offset: the case in which the
    classic kernel evaluates intermediates on k-extended boxes (or re-evaluates them per plane)."""
    names = {s["NAME"] for s in group["STENCILS"]}
    return any(box[2] != (0, 0) for s in group["STENCILS"] for f, box in s["OFFSETS"].items() if f in names)


class ColumnPlan:
    """Everything needed to render and launch the column kernel of one fused group."""

    column = True
    stream = False

    def __init__(self, program, group, opts, kernel_name=None):
        from .emitter import tile_geometry
        self.program, self.group, self.opts = program, group, opts
        self.id = group["ID"] + 1
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
        self.load = "tma"
        self.rj = max(1, int(opts["BLOCK"][0]))
        col = dict(opts.get("COLUMN") or {})
        warps = max(1, int(opts["THREADS"]) // 32)
        self.wx, self.wy = col.get("WARPS") or (2 if warps % 2 == 0 and warps > 1 else 1,
                                                 warps // 2 if warps % 2 == 0 and warps > 1 else warps)
        self.warps = self.wx * self.wy
        self.threads = self.warps * 32
        self.launch_threads = self.threads + 32
        self.kc = max(1, int(col.get("PLANES", 1)))
        self.unroll = max(1, int(col.get("UNROLL", max(1, 4 // self.kc))))   # ring slots per unrolled loop trip
        # registers are a per-scheduler resource (16384 per SM sub-partition, warps dealt round-robin): the cap
        # __launch_bounds__ yields is 16384 / (32 * ceil(warps / 4)) -- 255 up to 8 warps, 168 up to 12, 128 up
        # to 16, producer warp included.  MAXREG (__maxnreg__) only serves to lower it.
        self.maxreg = col.get("MAXREG")
        # WAVE_SYNC: producers start the n-th work item of every CTA together (see template_column.cu).  Measured:
        # the halo reads then hit L2 (dram reads 10.4 -> 8.2 GB on fused fast waves) but CTAs marching through the
        # same planes at the same time run at half the speed (5.4 vs 2.7 ms) -- off by default.
        self.wave_sync = bool(col.get("WAVE_SYNC", False))
        self.fuse_halo = bool(opts["FUSE_HALO"]) and all(s >= 2 * h for s, h in zip(dims, self.H))
        self.acc = {n: A.stencil_accesses(self.by_name[n]["LAMBDA"]) for n in self.names}
        self._schedule()
        if col.get("MERGE"):
            # The tile counts of the variant description are a cache-blocking choice of the model; results do not
            # depend on them.  A column kernel blocks for shared memory per plane, not per tile: with MERGE it
            # walks whole columns of its own shape -- the k-tiles of a column are merged (no pipeline refill per
            # k-tile) and x/y extents follow the warp grid instead of tiles narrower than a strip.
            self.N = (A.ceil_div(dims[0], self.wx * self.useful), A.ceil_div(dims[1], self.wy * self.rj), 1)
            geo = [tile_geometry(s, n) for s, n in zip(dims, self.N)]
            self.T = tuple(g[0] for g in geo)
            self.O = tuple(g[1] for g in geo)
            self.tiles = self.N[0] * self.N[1] * self.N[2]
        self._geometry(col.get("SLOTS"))
        self._cluster(col)
        self._arguments()

    # -- schedule ---------------------------------------------------------------------------
    def _schedule(self):
        names, acc, stencil = self.names, self.acc, self.by_name
        # how many planes ahead of the output plane every stencil runs
        lead = {}
        for s in reversed(names):
            if s not in lead:
                assert s in self.outputs, "stencil %s has no consumer and is not an output" % s
                lead[s] = 0
            if s in self.outputs:
                lead[s] = max(lead[s], 0)
            for f, (_, _, c) in acc[s]:
                if f in stencil:
                    assert names.index(f) < names.index(s), "stencil %s reads %s before it is computed" % (s, f)
                    lead[f] = max(lead[f], lead[s] + c) if f in lead else lead[s] + c
        self.lead = lead
        self.qtop = {}
        for s in names:
            for f, (_, _, c) in acc[s]:
                if f not in stencil:
                    self.qtop[f] = max(self.qtop.get(f, lead[s] + c), lead[s] + c)
        # thread-local rows on which every stencil is evaluated
        rows = {s: set(range(self.rj)) for s in self.outputs}
        for s in reversed(names):
            for f, (_, b, _) in acc[s]:
                if f in stencil:
                    rows.setdefault(f, set()).update(r + b for r in rows[s])
        self.rows = {s: list(range(min(rows[s]), max(rows[s]) + 1)) for s in names}
        # planes (relative to kk) at which every key = (field, column offset, row) is read
        uses = {}
        for s in names:
            for bb in self.rows[s]:
                for f, (a, b, c) in acc[s]:
                    uses.setdefault((f, a, bb + b), set()).add(lead[s] + c)
        for key in sorted(uses):
            f, a, bb = key
            if f in stencil and a != 0:
                uses.setdefault((f, 0, bb), set()).add(max(uses[key]))
        self.uses = uses

        def fresh_plane(key):
            f, a, _ = key
            if f not in stencil:
                return self.qtop[f]
            return lead[f] if a == 0 else max(uses[key])
        self.fresh_plane = fresh_plane

        # lanes of a warp whose value of a key is complete (shuffles lose one lane per column of reach)
        lanes = {}

        def lanes_of(key):
            if key in lanes:
                return lanes[key]
            f, a, bb = key
            if f not in stencil:
                out = (0, 31)
            elif a != 0:
                lo, hi = lanes_of((f, 0, bb))
                out = (max(0, lo - a), min(31, hi - a))
            else:
                lo, hi = 0, 31
                for g, (a2, b2, _) in acc[f]:
                    l2, h2 = lanes_of((g, a2, bb + b2))
                    lo, hi = max(lo, l2), min(hi, h2)
                out = (lo, hi)
            lanes[key] = out
            return out
        first, last = 0, 31
        for s in self.outputs:
            for bb in range(self.rj):
                lo, hi = lanes_of((s, 0, bb))
                first, last = max(first, lo), min(last, hi)
        assert first <= last, "shuffle chain longer than a warp"
        self.lanes_of, self.first, self.last = lanes_of, first, last
        self.useful = last - first + 1

        # the step program
        ops, fresh, declared = [], {}, []

        def name_of(key, q=None):
            f, a, bb = key
            base = "%s_%s_%s" % (f, _tag(a), _tag(bb))
            return base if q is None else "c_%s_q%s" % (base, _tag(q))

        def get_fresh(key):
            if key in fresh:
                return fresh[key]
            f, a, bb = key
            if f not in stencil:
                var = "l_" + name_of(key)
                ops.append(("load", var, f, a, bb))
            else:
                assert a != 0, "%s at row %d is read before it is evaluated" % (f, bb)
                var = "s_" + name_of(key)
                ops.append(("shfl", var, value((f, 0, bb), fresh_plane(key)), a))
            fresh[key] = var
            return var

        def value(key, q):
            top = fresh_plane(key)
            assert q <= top, "key %s wanted at plane %d above its newest plane %d" % (key, q, top)
            return get_fresh(key) if q == top else name_of(key, q)

        for s in names:
            for bb in self.rows[s]:
                operands = [value((f, a, bb + b), lead[s] + c) for f, (a, b, c) in acc[s]]
                var = "v_" + name_of((s, 0, bb))
                ops.append(("eval", var, s, bb, operands))
                fresh[(s, 0, bb)] = var
                if s in self.outputs and 0 <= bb < self.rj:
                    ops.append(("store", s, bb, var))
        carries = []
        for key in sorted(uses):
            top, low = fresh_plane(key), min(uses[key])
            if low < top:
                get_fresh(key)                      # the newest plane feeds the chain even if unused this step
                for q in range(low, top):
                    dst = name_of(key, q)
                    declared.append(dst)
                    carries.append((dst, fresh[key] if q + 1 == top else name_of(key, q + 1)))
        self.ops, self.carries, self.carried = ops, carries, declared
        # first step: every key is evaluated / loaded at its newest plane early enough for the lowest
        # plane anybody needs of it (relative to the first owned plane)
        kmin = 0
        for s in names:
            kmin = min(kmin, self.reach[s][2][0] - lead[s])
        for f in self.qtop:
            kmin = min(kmin, self.reach[f][2][0] - self.qtop[f])
        self.kmin = kmin

    # -- geometry and shared memory -----------------------------------------------------------
    def _geometry(self, slots):
        bx = min(self.T[0], self.wx * self.useful)
        by = min(self.T[1], self.wy * self.rj)
        self.B = (bx, by, self.T[2])
        self.nsub = (A.ceil_div(self.T[0], bx), A.ceil_div(self.T[1], by), 1)
        span_x, span_y = self.wx * self.useful, self.wy * self.rj
        cursor = 0
        self.input_plan = []
        for n, name in enumerate(self.inputs):
            keys = [k for k in self.uses if k[0] == name]
            amin, amax = min(k[1] for k in keys), max(k[1] for k in keys)
            bmin, bmax = min(k[2] for k in keys), max(k[2] for k in keys)
            # every lane of every strip reads staged cells, also the lanes left and right of the owning ones
            # (they evaluate the columns whose results the owning lanes take by shuffle): box column 0 is
            # column ``amin - first`` of the work item, the last strip ends 32 lanes after its first one
            dx = span_x - self.useful + 32 + amax - amin
            dx = dx + 1 + (dx + 1) % 2               # even start column + even extent (16-byte TMA rule)
            dy = span_y - self.rj + (bmax - bmin + 1)
            assert dx <= MAX_BOX and dy <= MAX_BOX, "column tile exceeds the TMA box limit"
            cells = dx * dy * self.kc
            self.input_plan.append({
                "NAME": name, "INDEX": n, "MAP": n, "AMIN": amin, "BMIN": bmin, "DIM": [dx, dy, self.kc],
                "XLO": amin - self.first,
                "ROW": dx, "PLANE": dx * dy, "CELLS": cells, "OFFSET": cursor, "QTOP": self.qtop[name],
            })
            cursor += _round_up(cells, ALIGN_DOUBLES)
        self.slot_doubles = cursor
        self.slot_bytes = sum(p["CELLS"] for p in self.input_plan) * 8
        budget = min(self.opts["SMEM_BUDGET"], SMEM_LIMIT)
        fixed = (FRONT_SLACK + TAIL_SLACK) * 8 + 128
        fit = (budget - fixed - 64) // (self.slot_doubles * 8 + 16)
        self.ns = int(slots) if slots else max(2, min(6, fit))
        self.stages = self.ns
        self.barrier_offset = (FRONT_SLACK + self.ns * self.slot_doubles + TAIL_SLACK) * 8
        self.smem_bytes = self.barrier_offset + 16 * self.ns + 48
        assert self.smem_bytes <= SMEM_LIMIT, "column ring (%d slots of %d B) exceeds shared memory" % (
            self.ns, self.slot_doubles * 8)

    def _cluster(self, col):
        """CLUSTER: (CX, CY) -- the CTAs of a thread-block cluster take the CX x CY tiles of one *super-tile* and
        their TMA producers stay within SLACK ring slots of each other (one remote ``mbarrier`` arrive per peer and
        slot, template_column.cu).  Neighbouring columns stage overlapping boxes -- the halo lanes and rows, the
        16-byte rule, partly used DRAM lines: 1.56x the input bytes on fused fast waves -- and persistent CTAs that
        are left alone drift apart by more planes than the L2 holds, so every CTA fetched its halo from HBM.  Kept
        within a few planes of each other, the CTAs of a cluster find each other's halo in L2.  Needs work items
        that are whole columns of one shape (MERGE), full-domain programs (not the training loop)."""
        want = col.get("CLUSTER") or (1, 1)
        cx, cy = (int(want[0]), int(want[1])) if not isinstance(want, int) else (int(want), 1)
        cx, cy = max(1, min(cx, self.N[0])), max(1, min(cy, self.N[1]))
        usable = (cx * cy > 1 and cx * cy <= 8 and self.nsub == (1, 1, 1) and self.N[2] == 1
                  and not self.program.get("TRAINING") and not self.wave_sync)
        self.cluster = (cx, cy) if usable else (1, 1)
        self.cluster_size = self.cluster[0] * self.cluster[1]
        self.slack = max(1, int(col.get("SLACK", 4)))          # ring slots a producer may run ahead of its peers
        self.super = (A.ceil_div(self.N[0], self.cluster[0]), A.ceil_div(self.N[1], self.cluster[1]))
        # every work item of a cluster kernel has the same number of ring slots (the tick protocol counts them)
        self.item_chunks = A.ceil_div(self.S[2] - self.kmin, self.kc)
        if self.cluster_size > 1:
            ticks = 2 * self.slack + 1
            self.tick_offset = self.smem_bytes - 48 + 16       # behind full[] / empty[]
            self.smem_bytes = self.tick_offset + 8 * ticks + 48
            assert self.smem_bytes <= SMEM_LIMIT, "column ring plus cluster barriers exceed shared memory"

    def _arguments(self):
        off = 0
        self.args = []
        for p in self.input_plan:
            self.args.append(("tmap", off, p["NAME"], p["DIM"]))
            off += 128
        for name in self.inputs:
            self.args.append(("ptr", off, name, None))
            off += 8
        for name in self.outputs:
            self.args.append(("ptr", off, name, None))
            off += 8
        self.args.append(("tilectl", off, None, None))
        off += 8
        if self.fuse_halo:
            self.args.append(("remote", off, None, None))
            off += 16
        self.scratch_bytes = 0
        if self.wave_sync and self.nsub == (1, 1, 1):
            self.args.append(("scratch", off, None, None))       # wave counters: SYNC_LINES 128-byte lines per CTA
            off += 8
            self.scratch_bytes = 128 * SYNC_LINES
        else:
            self.wave_sync = False
        self.params_bytes = _round_up(off, 64)

    # -- CUDA text --------------------------------------------------------------------------
    def _body(self):
        """Straight-line code of one k-step for one thread (the IR rendered as C)."""
        from .emitter import branchify
        plan = {p["NAME"]: p for p in self.input_plan}
        lines = []
        for op in self.ops:
            if op[0] == "load":
                _, var, f, a, bb = op
                p = plan[f]
                lines.append("const double %s = pl_%s[%d];" % (var, f, (bb - p["BMIN"]) * p["ROW"] + a - p["AMIN"]))
            elif op[0] == "shfl":
                _, var, src, a = op
                fn = "__shfl_down_sync" if a > 0 else "__shfl_up_sync"
                lines.append("const double %s = %s(0xffffffffu, %s, %d);" % (var, fn, src, abs(a)))
            elif op[0] == "eval":
                _, var, s, bb, operands = op
                it = iter(operands)
                body = A._ACCESS.sub(lambda m: next(it), self.by_name[s]["LAMBDA"])
                if self.opts["BRANCH"]:
                    body = branchify(body)
                lines.append("double %s; { %s %s = res; }" % (var, body, var))
            else:
                _, s, bb, var = op
                store = ("__stcs(&o_%s[%d * ROW], %s);" if self.opts["STORE_CS"] else "o_%s[%d * ROW] = %s;") % (s, bb, var)
                if self.fuse_halo:
                    store += (" if (w.edge || p_%s < HZ || p_%s >= Z - HZ) store_wrapped(o_%s, %d * (long)ROW, "
                              "w.oi + col, w.oj + row0 + %d, p_%s, %s, %d, prm.remote[0], prm.remote[1]);"
                              % (s, s, s, bb, bb, s, var, self.outputs.index(s)))
                lines.append("if (ok_%d && pk_%s) { %s }" % (bb, s, store))
        for dst, src in self.carries:
            lines.append("%s = %s;" % (dst, src))
        return lines

    def context(self):
        return {
            "COLUMN": True, "STREAM": False, "ID": self.id, "KERNEL": self.kernel_name,
            "THREADS": self.threads, "WARPS": self.warps, "LAUNCH_THREADS": self.launch_threads,
            "WX": self.wx, "WY": self.wy, "RJ": self.rj, "KC": self.kc, "WAVE_SYNC": self.wave_sync, "SYNC_LINES": SYNC_LINES, "UNROLL": self.unroll, "MAXREG": self.maxreg, "NS": self.ns, "STAGES": self.ns,
            "FIRST": self.first, "LAST": self.last, "USEFUL": self.useful, "KMIN": self.kmin,
            "T": self.T, "O": self.O, "N": self.N, "B": self.B, "NSUB": self.nsub,
            "NSUB_TOTAL": self.nsub[0] * self.nsub[1], "TILES": self.tiles, "TMA": True,
            "INPUTS": self.input_plan, "NIN": len(self.inputs), "NMAPS": len(self.input_plan),
            "OUTPUTS": self.outputs, "NOUT": len(self.outputs),
            "OUT_LEADS": [{"NAME": n, "INDEX": self.outputs.index(n), "LEAD": self.lead[n]} for n in self.outputs],
            "SLOT_DOUBLES": self.slot_doubles, "SLOT_BYTES": self.slot_bytes,
            "BARRIER_OFFSET": self.barrier_offset, "SMEM_BYTES": self.smem_bytes,
            "PARAMS_BYTES": self.params_bytes, "NAMES": self.names, "FUSE_HALO": self.fuse_halo,
            "FRONT": FRONT_SLACK, "CARRIED": self.carried, "BODY": self._body(),
            "STREAM_STORES": bool(self.opts["STORE_CS"]), "LEADS": self.lead,
            "CLUSTER": self.cluster_size if self.cluster_size > 1 else 0, "CX": self.cluster[0], "CY": self.cluster[1],
            "SNX": self.super[0], "SN_TOTAL": self.super[0] * self.super[1], "SLACK": self.slack,
            "TICKS": 2 * self.slack + 1, "TICK_OFFSET": getattr(self, "tick_offset", 0),
            "ITEM_CHUNKS": self.item_chunks,
        }

    # -- verification on symbolic tags ---------------------------------------------------------
    def simulate(self, planes=None):
        """Replay the step program of one warp on tags and check every store.

        A tag says what a register holds: ``(field, column, row, plane)``.  Loads produce the tag of
        the staged cell they address, an evaluation checks that each operand is exactly the cell its
        access names and yields the stencil's own tag, shuffles move tags between lanes, carries move
        them between steps.  Registers start poisoned; a store inside the owned planes must see the
        right tag on every owning lane.  Returns the number of stores checked.
        """
        planes = planes or (4 - self.kmin)
        poison = None
        boxes = {p["NAME"]: p for p in self.input_plan}
        state = {name: [poison] * 32 for name in self.carried}
        checked = 0
        for step in range(planes - self.kmin):
            kk = self.kmin + step                      # owned planes are 0 .. planes-1
            for op in self.ops:
                if op[0] == "load":
                    _, var, f, a, bb = op
                    p = boxes[f]
                    # staged cell of the last strip / last warp row, worst-case parity: must be inside the box
                    inside = [0 <= lane + a - p["AMIN"] and (self.wx - 1) * self.useful + lane + a - p["AMIN"] + 1
                              < p["DIM"][0] and 0 <= bb - p["BMIN"] and (self.wy - 1) * self.rj + bb - p["BMIN"]
                              < p["DIM"][1] for lane in range(32)]
                    state[var] = [(f, lane - self.first + a, bb, kk + self.qtop[f]) if inside[lane] else poison
                                  for lane in range(32)]
                elif op[0] == "shfl":
                    _, var, src, a = op
                    state[var] = [state[src][lane + a] if 0 <= lane + a < 32 else poison for lane in range(32)]
                elif op[0] == "eval":
                    _, var, s, bb, operands = op
                    out = []
                    for lane in range(32):
                        col, ok = lane - self.first, True
                        for (f, (a, b, c)), operand in zip(self.acc[s], operands):
                            have = state[operand][lane]
                            if have is poison:
                                ok = False
                            else:
                                want = (f, col + a, bb + b, kk + self.lead[s] + c)
                                assert have == want, "%s row %d lane %d step %d: operand %s holds %s, wanted %s" % (
                                    s, bb, lane, step, operand, have, want)
                        out.append((s, col, bb, kk + self.lead[s]) if ok else poison)
                    state[var] = out
                else:
                    _, s, bb, var = op
                    plane = kk + self.lead[s]
                    if 0 <= plane < planes:
                        for lane in range(self.first, self.last + 1):
                            want = (s, lane - self.first, bb, plane)
                            assert state[var][lane] == want, "store of %s row %d lane %d plane %d sees %s" % (
                                s, bb, lane, plane, state[var][lane])
                            checked += 1
            new = {dst: state[src] for dst, src in self.carries}
            # carries are emitted low plane first, so sequential assignment equals this parallel one
            order = [dst for dst, _ in self.carries]
            for n, (dst, src) in enumerate(self.carries):
                assert src not in order[:n], "carry %s <- %s reads a register already overwritten" % (dst, src)
            state.update(new)
        return checked
