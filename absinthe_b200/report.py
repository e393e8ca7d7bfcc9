"""Evaluation of an experiment folder: estimates vs measurements (ref diffusion.R, advection.R, fastwaves.R).

The reference evaluates a run with R scripts: ``summarySE`` (ref SPCL_Stats.R:24-99) condenses the rows
of ``results.csv`` per variant (measured time = TOTAL - HALO, ref diffusion.R:17-19), the table is merged
with ``estimates.csv`` and drawn as an estimate/measurement scatter plot with the OPT / HAND / AUTO / MIN /
MAX variants labelled (ref diffusion.R:33-52); the auto-tuning table ``auto.csv`` is reduced per group to
the variants whose lower confidence bound reaches the group's best median (ref diffusion.R:84-108).  R is
not part of this image, so the same numbers are computed here with numpy and the plot is written as SVG.

    python -m absinthe_b200.report ./diffusion            # summary.csv, scatter.svg, printed report
"""

import csv
import math
import os
import sys
from statistics import NormalDist

import numpy as np

from .fit import spearman

LABELLED = ("OPT", "HAND", "AUTO", "MIN", "MAX")


def _student_t_quantile(p, df):
    from scipy.stats import t
    return float(t.ppf(p, df))


def rank_interval(values, conf=0.95):
    """Distribution-free confidence interval of the median (ref SPCL_Stats.R:34-59).

    Ranks are R's (1-based) ``floor((n - z sqrt n) / 2)`` and ``ceiling(1 + (n + z sqrt n) / 2)``; a rank
    outside ``1..n`` indexes nothing in R and yields NA -- ``nan`` here.
    """
    data = np.sort(np.asarray(values, dtype=float))
    n = len(data)
    z = -NormalDist().inv_cdf((1.0 - conf) / 2.0)
    low = math.floor((n - z * math.sqrt(n)) / 2.0)
    high = math.ceil(1.0 + (n + z * math.sqrt(n)) / 2.0)

    def pick(rank):
        return float(data[rank - 1]) if 1 <= rank <= n else float("nan")
    return pick(low), pick(high)


def summary_se(values, conf=0.95, quantile_interval=0.95):
    """One ``summarySE`` row: the columns of ref SPCL_Stats.R:66-97 for one group of measurements."""
    v = np.asarray(values, dtype=float)
    n = len(v)
    sd = float(v.std(ddof=1)) if n > 1 else float("nan")
    se = sd / math.sqrt(n) if n else float("nan")
    q = (1.0 - quantile_interval) / 2.0
    cil, cih = rank_interval(v, conf)
    return {
        "NumMeasures": n, "min": float(v.min()), "max": float(v.max()), "mean": float(v.mean()),
        "median": float(np.median(v)), "StandardDev": sd,
        "Quantile.low": float(np.quantile(v, q)), "Quantile.high": float(np.quantile(v, 1.0 - q)),
        "cil": cil, "cih": cih, "StandardErr": se,
        "CI.Norm.StudT": se * _student_t_quantile(conf / 2.0 + 0.5, n - 1) if n > 1 else float("nan"),
        "CI.Norm.Norm": se * NormalDist().inv_cdf(conf / 2.0 + 0.5),
    }


def read_rows(path):
    with open(path, newline="") as handle:
        return list(csv.DictReader(handle))


def measured(rows):
    """VAR -> list of TOTAL - HALO [ms], in file order (ref diffusion.R:17-19)."""
    out = {}
    for row in rows:
        out.setdefault(row["VAR"], []).append(float(row["TOTAL"]) - float(row["HALO"]))
    return out


def merge(estimates, results, conf=0.95):
    """Rows (sorted by VAR, as R's ``merge`` does) of the variants present in both tables."""
    est = {row["VAR"]: float(row["EST"]) for row in estimates}
    table = []
    for var, values in sorted(measured(results).items()):
        if var in est:
            row = {"VAR": var, "EST": est[var]}
            row.update(summary_se(values, conf))
            table.append(row)
    return table


def evaluate(table):
    """The printed part of the R scripts plus the rank correlation the north star asks for."""
    by_name = {row["VAR"]: row for row in table}
    report = {"variants": len(table),
              "spearman": spearman([r["EST"] for r in table], [r["median"] for r in table]) if len(table) > 2
              else float("nan")}
    opt, auto = by_name.get("OPT"), by_name.get("AUTO")
    if opt:
        # the measurements that beat the optimal solution (ref diffusion.R:54-56)
        report["better"] = [r["VAR"] for r in table if r["median"] <= opt["median"] + 0.01]
    if opt and auto:
        report["auto_vs_opt_percent"] = (auto["median"] / opt["median"] - 1.0) * 100.0   # ref diffusion.R:31
        report["accuracy"] = 1.0 - auto["median"] / opt["median"]                        # ref diffusion.R:59-60
    return report


def auto_tuning(rows, conf=0.95):
    """Per group: best median and the variants whose lower bound reaches it (ref diffusion.R:84-108).

    The group is the first number of the variant name (ref diffusion.R:88-92).
    """
    import re
    groups = {}
    for var, values in measured(rows).items():
        found = re.findall(r"[0-9]+", var)
        row = {"VAR": var}
        row.update(summary_se(values, conf))
        groups.setdefault(found[0] if found else "", []).append(row)
    out = {}
    for grp, members in groups.items():
        optimum = min(m["median"] for m in members)
        out[grp] = {"optimum": optimum,
                    "candidates": [m for m in members if not (m["cil"] > optimum)]}
    return out


def write_summary(path, table):
    columns = ["VAR", "EST", "NumMeasures", "min", "max", "mean", "median", "StandardDev", "Quantile.low",
               "Quantile.high", "cil", "cih", "StandardErr", "CI.Norm.StudT", "CI.Norm.Norm"]
    with open(path, "w", newline="") as handle:
        writer = csv.DictWriter(handle, fieldnames=columns)
        writer.writeheader()
        writer.writerows(table)


def scatter_svg(path, table, title, size=360, margin=48):
    """Estimate (x) against median measurement (y) with the cil..cih bars and the y = x line."""
    pts = [r for r in table if math.isfinite(r["EST"]) and math.isfinite(r["median"])]
    if not pts:
        return
    lo = min(min(r["EST"], r["median"]) for r in pts)
    hi = max(max(r["EST"], r["median"]) for r in pts)
    pad = 0.05 * (hi - lo or 1.0)
    lo, hi = lo - pad, hi + pad

    def sx(v):
        return margin + (v - lo) / (hi - lo) * (size - 2 * margin)

    def sy(v):
        return size - margin - (v - lo) / (hi - lo) * (size - 2 * margin)
    out = ['<svg xmlns="http://www.w3.org/2000/svg" width="%d" height="%d" font-family="sans-serif" font-size="10">'
           % (size, size),
           '<rect width="100%" height="100%" fill="white"/>',
           '<rect x="%d" y="%d" width="%d" height="%d" fill="none" stroke="black"/>'
           % (margin, margin, size - 2 * margin, size - 2 * margin),
           '<line x1="%.1f" y1="%.1f" x2="%.1f" y2="%.1f" stroke="black"/>' % (sx(lo), sy(lo), sx(hi), sy(hi)),
           '<text x="%d" y="16" text-anchor="middle" font-size="12">%s</text>' % (size // 2, title),
           '<text x="%d" y="%d" text-anchor="middle">estimated time [ms]</text>' % (size // 2, size - 8),
           '<text transform="translate(12 %d) rotate(-90)" text-anchor="middle">measured time [ms]</text>' % (size // 2)]
    for frac in (0.0, 0.5, 1.0):
        v = lo + frac * (hi - lo)
        out.append('<text x="%.1f" y="%d" text-anchor="middle">%.3g</text>' % (sx(v), size - margin + 12, v))
        out.append('<text x="%d" y="%.1f" text-anchor="end">%.3g</text>' % (margin - 4, sy(v) + 3, v))
    for r in sorted(pts, key=lambda r: r["VAR"] in LABELLED):        # labelled variants drawn on top
        x, y = sx(r["EST"]), sy(r["median"])
        if math.isfinite(r["cil"]) and math.isfinite(r["cih"]):
            out.append('<line x1="%.1f" y1="%.1f" x2="%.1f" y2="%.1f" stroke="#999"/>' % (x, sy(r["cil"]), x, sy(r["cih"])))
        if r["VAR"] in LABELLED:
            out.append('<rect x="%.1f" y="%.1f" width="6" height="6" fill="#2b8cbe"/>' % (x - 3, y - 3))
            out.append('<text x="%.1f" y="%.1f">%s</text>' % (x + 5, y - 4, r["VAR"].lower()))
        else:
            out.append('<circle cx="%.1f" cy="%.1f" r="2.4" fill="#bae4bc" stroke="#7bccc4"/>' % (x, y))
    out.append("</svg>")
    with open(path, "w") as handle:
        handle.write("\n".join(out) + "\n")


def main(argv=None):
    import argparse
    import json
    ap = argparse.ArgumentParser(description="evaluate estimates.csv / results.csv / auto.csv of an experiment folder")
    ap.add_argument("folder")
    ap.add_argument("--conf", type=float, default=0.95)
    args = ap.parse_args(argv)
    name = os.path.basename(os.path.normpath(args.folder))
    est, res, auto = (os.path.join(args.folder, f) for f in ("estimates.csv", "results.csv", "auto.csv"))
    if os.path.exists(est) and os.path.exists(res):
        table = merge(read_rows(est), read_rows(res), args.conf)
        write_summary(os.path.join(args.folder, "summary.csv"), table)
        scatter_svg(os.path.join(args.folder, "scatter.svg"), table, name)
        print(json.dumps(evaluate(table), indent=1))
    if os.path.exists(auto):
        print("auto tuning")
        for grp, info in sorted(auto_tuning(read_rows(auto), args.conf).items()):
            print(info["optimum"])
            for row in info["candidates"]:
                print("  %s %s median %.6g cil %.6g cih %.6g" % (row["VAR"], grp, row["median"], row["cil"], row["cih"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
