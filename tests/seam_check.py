"""Parity of the *bench* variants at full size, single- or multi-GPU, against the CPU oracle on windows.

Test infrastructure (it imports ``oracle/``): used by ``bench.py --check`` and ``tests/test_distributed.py``.
The inputs are an exact integer hash of the *global* periodic coordinates, so every rank can build its own
slab (halos included) and the oracle's window inputs -- also the part of a window that lies in a neighbour's
slab -- without any communication.  One step of the variants runs exactly like the timed step of
``bench.py`` (group kernels, fused periodic copies, ``SlabExchange`` when N > 1); then windows at the domain
corner, in the middle and across both j-seams of the slab are compared bit for bit with the oracle wherever
the window's own periodic wrap cannot reach (``MARGIN``).
"""

import numpy as np

MARGIN = 6          # > dependency radius of either sequence
WINDOW = (40, 28, 14)
H = 3


from absinthe_b200.fields import field_seed, global_field  # noqa: E402,F401  (re-exported for the tests)


def run_check(programs, world=1, rank=0, device=0, windows=None, exchange_kind="p2p"):
    """``programs``: [(kernel, program)] as ``bench.workload_programs()`` yields them.  Returns a report dict with
    ``ok`` and one entry per (kernel, window)."""
    import torch
    from absinthe_b200 import programs as P
    from absinthe_b200.distributed import P2PExchange, SlabExchange
    from absinthe_b200.launcher import Variant
    from oracle.reference_oracle import Oracle
    report = {"ok": True, "rank": rank, "world": world, "windows": []}
    dev = "cuda:%d" % device
    for kernel, program in programs:
        v = Variant.from_program(P.clone(program), device=device)
        X, Y, Z = v.X, v.Y, v.Z
        extent = (X, world * Y, Z)
        j_first = rank * Y                                  # global row of the slab's first interior row
        mine = ((-H, X + H), (j_first - H, j_first + Y + H), (-H, Z + H))
        ins = [global_field(field_seed(kernel, n), mine, extent, dev) for n in v.inputs]
        outs = [v.empty_field() for _ in v.outputs]
        v.bind(ins, outs)
        exchange = (P2PExchange if exchange_kind == "p2p" else SlabExchange)(v, world, rank) if world > 1 else None
        for g in range(1, v.n_groups + 1):                  # one step, as bench.py runs it
            if exchange is not None and g > 1:
                exchange.wait(g - 1)
            if exchange is not None:
                exchange.apply(g)
            else:
                v.apply_group(g, with_halo=False)
                v.halo_group(g)
        if exchange is not None:
            for g in range(1, v.n_groups + 1):
                exchange.wait(g)
        torch.cuda.synchronize(device)
        wx, wy, wz = WINDOW
        # (ox, oy, oz): window origins in local interior coordinates; the two seam windows straddle the slab's edges
        spots = windows or [(0, 0, 0), (X - wx, Y // 2, Z - wz), (X // 2 - 21, -wy // 2, 29), (7, Y - wy // 2, 0)]
        small = P.make_program(kernel, domain=WINDOW)
        oracle = Oracle(small)
        for ox, oy, oz in spots:
            box = ((ox - H, ox + wx + H), (j_first + oy - H, j_first + oy + wy + H), (oz - H, oz + wz + H))
            win = [global_field(field_seed(kernel, n), box, extent, dev).cpu().numpy() for n in oracle.inputs]
            want = oracle.run(win)
            # rows of the window this rank holds (interior or halo), inside the window's trustworthy core
            rows = [pj for pj in range(H + MARGIN, H + wy - MARGIN) if 0 <= oy + pj < Y + 2 * H]
            entry = {"kernel": kernel, "window": [ox, oy, oz], "rows": len(rows), "ok": True}
            if rows:
                lo, hi = rows[0], rows[-1] + 1
                for name, w in zip(oracle.outputs, want):
                    have = outs[v.outputs.index(name)][oz + H + MARGIN:oz + H + wz - MARGIN, oy + lo:oy + hi,
                                                       ox + H + MARGIN:ox + H + wx - MARGIN].cpu().numpy()
                    core = w[H + MARGIN:H + wz - MARGIN, lo:hi, H + MARGIN:H + wx - MARGIN]
                    if not np.array_equal(have, core):
                        entry["ok"] = report["ok"] = False
                        entry.setdefault("differs", []).append(name)
            report["windows"].append(entry)
        v.close()
        del ins, outs
        torch.cuda.empty_cache()
    return report
