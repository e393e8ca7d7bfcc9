"""Re-training of the analytical model's machine parameters from training-sweep timings.

Python replacement of the regression part of the reference's ``fitcache.R`` / ``fitddr.R``
(SURVEY.md section 8 f-2): least-absolute-deviations fit of the measured time of a training program on
per-tile regressors, solved as a linear program (``scipy.optimize.linprog``, HiGHS) instead of R's
``L1pack::lad``.  On the reference's shipped ``fitcache/results.csv`` / ``fitddr/results.csv`` it
reproduces the parameters printed in the reference README (tests/test_fit.py).

GPU reading of the parameters: a tile is a CTA's unit of work, a core an SM,
``CACHE BODY/PEEL`` the per-fetch / per-row cost of on-chip (shared memory, registers, L1/L2) operand
traffic, ``MEMORY RW/ST BODY/PEEL`` the per-cell / per-row cost of HBM streams.
"""

import csv

import numpy as np

STEPS = 9           # stencils per training program (ref fitcache.R:27, fitddr.R:27)


def load_results(path):
    """``results.csv`` -> list of (VAR, X, Y, Z, TOTAL) with numeric columns converted."""
    rows = []
    with open(path) as handle:
        for rec in csv.DictReader(handle, quotechar="'"):
            rows.append((rec["VAR"], int(rec["X"]), int(rec["Y"]), int(rec["Z"]), float(rec["TOTAL"])))
    return rows


def _tile_terms(rows, cores):
    """Per-tile extents of every sample (ref fitcache.R:19-25): the domain is 2 x 2 x 5*cores tiles."""
    X = np.array([r[1] for r in rows], dtype=float)
    Y = np.array([r[2] for r in rows], dtype=float)
    Z = np.array([r[3] for r in rows], dtype=float)
    return {"XYZ": X * Y * Z / (2 * 2 * 5 * cores), "XY": X * Y / 4, "XZ": X * Z / (2 * 5 * cores),
            "YZ": Y * Z / (2 * 5 * cores), "X": X / 2, "Y": Y / 2, "Z": Z / (5 * cores)}


def lad(design, target, nonnegative=False):
    """argmin_b sum |design @ b - target|  via  min 1'(u+v)  s.t.  design b + u - v = target."""
    from scipy.optimize import linprog
    from scipy.sparse import eye, hstack, csr_matrix
    n, p = design.shape
    cost = np.concatenate([np.zeros(p), np.ones(2 * n)])
    a_eq = hstack([csr_matrix(design), eye(n), -eye(n)], format="csr")
    bounds = [(0 if nonnegative else None, None)] * p + [(0, None)] * (2 * n)
    # scale columns so HiGHS sees O(1) coefficients
    scale = np.abs(design).max(axis=0)
    scale[scale == 0] = 1.0
    a_eq = hstack([csr_matrix(design / scale), eye(n), -eye(n)], format="csr")
    res = linprog(cost, A_eq=a_eq, b_eq=target, bounds=bounds, method="highs")
    if not res.success:
        raise RuntimeError("LAD fit failed: " + res.message)
    return res.x[:p] / scale


def r_squared(design, target, coef):
    pred = design @ coef
    return 1.0 - float(np.sum((target - pred) ** 2) / np.sum((target - target.mean()) ** 2))


def _with_intercept(design, intercept):
    return np.concatenate([design, np.ones((len(design), 1))], axis=1) if intercept else design


def fit_cache(rows, cores, nonnegative=False, intercept=False, tiles_per_core=1):
    """``{"BODY", "PEEL"}`` from a fitcache sweep (ref fitcache.R:27-50; PT8 is excluded, ``fac >= 10``).

    ``intercept`` adds a constant column: on a GPU one wave of small tiles is dominated by the launch
    and pipeline-fill latency, which the reference's through-the-origin model cannot express.
    ``tiles_per_core``: tiles every core executed per run (20 / STRIDE of the training programs); the
    measured time is divided by it, so the regressors stay per-tile as in the reference."""
    fac = np.array([float(r[0][2:]) for r in rows])
    keep = fac >= 10
    rows = [r for r, k in zip(rows, keep) if k]
    fac = fac[keep]
    t = _tile_terms(rows, cores)
    design = _with_intercept(np.stack([fac * STEPS * t["XYZ"], fac * STEPS * t["YZ"]], axis=1), intercept)
    target = np.array([r[4] for r in rows]) / tiles_per_core
    coef = lad(design, target, nonnegative)
    out = {"BODY": coef[0] / cores, "PEEL": coef[1] / cores, "R2": r_squared(design, target, coef),
           "SAMPLES": len(rows)}
    if intercept:
        out["LAUNCH"] = coef[2]
    return out


def fit_ddr(rows, cores, nonnegative=False, intercept=False, tiles_per_core=1):
    """``{"RW BODY", "ST BODY", "RW PEEL", "ST PEEL"}`` from a fitddr sweep (ref fitddr.R:27-56)."""
    inp = np.array([float(r[0][2]) for r in rows])
    keep = inp >= 1
    rows = [r for r, k in zip(rows, keep) if k]
    inp = inp[keep]
    bd = np.array([float(r[0][5]) * 2 for r in rows])
    t = _tile_terms(rows, cores)
    rw_body = STEPS * t["XYZ"]
    st_body = STEPS * (inp + 1) * t["XYZ"] + STEPS * inp * bd * (t["XY"] + t["XZ"] + t["YZ"])
    rw_peel = STEPS * t["YZ"]
    st_peel = STEPS * (inp + 1) * t["YZ"] + STEPS * inp * bd * (t["Y"] + t["Z"])
    design = _with_intercept(np.stack([rw_body, st_body, rw_peel, st_peel], axis=1), intercept)
    target = np.array([r[4] for r in rows]) / tiles_per_core
    coef = lad(design, target, nonnegative)
    names = ("RW BODY", "ST BODY", "RW PEEL", "ST PEEL")
    out = {n: c / cores for n, c in zip(names, coef)}
    out.update(R2=r_squared(design, target, coef), SAMPLES=len(rows))
    if intercept:
        out["LAUNCH"] = coef[4]
    return out


def machine_model(cache, memory, cores, capacity):
    """The three dict entries of a driver ``PROGRAM`` that describe the machine (ref diffusion.py:71-73)."""
    return {"MACHINE": {"CORES": cores, "CAPACITY": capacity},
            "MEMORY": {k: memory[k] for k in ("RW BODY", "ST BODY", "RW PEEL", "ST PEEL")},
            "CACHE": {k: cache[k] for k in ("BODY", "PEEL")}}


def spearman(a, b):
    """Rank correlation (average ranks for ties) -- the model-quality figure of the north star."""
    def ranks(v):
        v = np.asarray(v, dtype=float)
        order = np.argsort(v, kind="mergesort")
        r = np.empty(len(v))
        r[order] = np.arange(len(v), dtype=float)
        for value in np.unique(v):
            idx = np.where(v == value)[0]
            r[idx] = r[idx].mean()
        return r
    ra, rb = ranks(a), ranks(b)
    ra -= ra.mean()
    rb -= rb.mean()
    return float(ra @ rb / np.sqrt((ra @ ra) * (rb @ rb)))


def main(argv=None):
    import argparse
    import json
    ap = argparse.ArgumentParser(description="fit the model parameters from training results")
    ap.add_argument("--cache", help="fitcache results.csv")
    ap.add_argument("--ddr", help="fitddr results.csv")
    ap.add_argument("--cores", type=int, default=148)
    ap.add_argument("--nonnegative", action="store_true")
    ap.add_argument("--intercept", action="store_true", help="fit a constant launch-latency term as well")
    ap.add_argument("--tiles-per-core", type=float, default=1.0,
                    help="tiles each core executed per run: 20 / the --stride the sweep was generated with")
    args = ap.parse_args(argv)
    out = {}
    if args.cache:
        out["CACHE"] = fit_cache(load_results(args.cache), args.cores, args.nonnegative, args.intercept,
                                 args.tiles_per_core)
    if args.ddr:
        out["MEMORY"] = fit_ddr(load_results(args.ddr), args.cores, args.nonnegative, args.intercept,
                                args.tiles_per_core)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
