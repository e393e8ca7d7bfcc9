"""Evaluation of estimates vs measurements (absinthe_b200.report; ref diffusion.R, SPCL_Stats.R)."""

import csv
import json
import math
import os

import numpy as np
import pytest

from absinthe_b200 import report as R
from absinthe_b200.fit import spearman

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_rank_interval_follows_the_r_ranks():
    # n = 48 (64 runs, 16 skipped): z sqrt(n) = 13.58 -> ranks floor(17.2) = 17 and ceil(31.79) = 32
    values = list(range(48, 0, -1))
    assert R.rank_interval(values, 0.95) == (17.0, 32.0)
    # n = 64 (auto-tuning, nothing skipped): ranks floor(24.16) = 24, ceil(40.84) = 41
    assert R.rank_interval(range(1, 65), 0.95) == (24.0, 41.0)
    # too few measurements: R indexes rank 0 / n + 1 and gets NA
    low, high = R.rank_interval([1.0, 2.0, 3.0], 0.95)
    assert math.isnan(low) and math.isnan(high)


def test_summary_se_columns():
    rng = np.random.default_rng(7)
    v = rng.normal(2.0, 0.1, 48)
    s = R.summary_se(v)
    assert s["NumMeasures"] == 48
    assert s["mean"] == pytest.approx(v.mean())
    assert s["median"] == pytest.approx(np.median(v))
    assert s["StandardDev"] == pytest.approx(v.std(ddof=1))
    assert s["StandardErr"] == pytest.approx(v.std(ddof=1) / math.sqrt(48))
    assert s["Quantile.low"] == pytest.approx(np.quantile(v, 0.025))       # R's default (type 7) quantile
    assert s["Quantile.high"] == pytest.approx(np.quantile(v, 0.975))
    assert s["CI.Norm.Norm"] == pytest.approx(1.959964 * s["StandardErr"], rel=1e-6)
    assert s["CI.Norm.StudT"] == pytest.approx(2.011741 * s["StandardErr"], rel=1e-6)    # qt(0.975, 47)
    assert s["cil"] <= s["median"] <= s["cih"]


def test_spearman_of_the_shipped_tables_matches_the_baseline():
    """Spearman(estimate, measured) = 0.475 / 0.033 / 0.591 over the shipped variants."""
    shipped = json.load(open(os.path.join(GOLDEN, "shipped_medians.json")))
    expected = {"diffusion": (132, 0.475), "advection": (124, 0.033), "fastwaves": (137, 0.591)}
    for kernel, (count, rho) in expected.items():
        rows = list(shipped[kernel].values())
        assert len(rows) == count
        assert spearman([r[0] for r in rows], [r[1] for r in rows]) == pytest.approx(rho, abs=6e-4)


def write_csv(path, header, rows):
    with open(path, "w", newline="") as handle:
        writer = csv.writer(handle)
        writer.writerow(header)
        writer.writerows(rows)


def test_folder_evaluation(tmp_path, capsys):
    rng = np.random.default_rng(3)
    true = {"OPT": 1.0, "AUTO": 1.1, "HAND": 1.4, "MIN": 2.0, "MAX": 1.7, "k-0-0-1": 0.99, "k-0-1-1": 1.6}
    write_csv(tmp_path / "estimates.csv", ["VAR", "EST"], [(k, v * 1.05) for k, v in true.items()] + [("unmeasured", 3)])
    rows = []
    for var, t in true.items():
        for _ in range(48):
            halo = 0.1 + 0.01 * rng.random()
            rows.append((var, 64, 64, 60, t + halo + 0.001 * rng.random(), halo))
    write_csv(tmp_path / "results.csv", ["VAR", "X", "Y", "Z", "TOTAL", "HALO"], rows)
    auto = []
    for var, t in {"k-0-1x1x60": 1.0, "k-0-1x1x30": 1.0005, "k-0-2x2x60": 1.3, "k-1-1x1x60": 0.5}.items():
        for _ in range(64):
            auto.append((var, 64, 64, 60, t + 0.001 * rng.random(), 0.0))
    write_csv(tmp_path / "auto.csv", ["VAR", "X", "Y", "Z", "TOTAL", "HALO"], auto)

    table = R.merge(R.read_rows(tmp_path / "estimates.csv"), R.read_rows(tmp_path / "results.csv"))
    assert [r["VAR"] for r in table] == sorted(true)           # merged on VAR, unmeasured variants dropped
    by = {r["VAR"]: r for r in table}
    assert by["MIN"]["median"] == pytest.approx(2.0005, abs=1e-3)       # TOTAL - HALO
    ev = R.evaluate(table)
    assert ev["variants"] == 7 and ev["spearman"] == pytest.approx(1.0)
    assert ev["better"] == ["OPT", "k-0-0-1"]
    assert ev["auto_vs_opt_percent"] == pytest.approx(10.0, abs=0.1)
    assert ev["accuracy"] == pytest.approx(-0.1, abs=1e-3)

    tuning = R.auto_tuning(R.read_rows(tmp_path / "auto.csv"))
    assert sorted(tuning) == ["0", "1"]
    assert tuning["0"]["optimum"] == pytest.approx(1.0005, abs=1e-3)
    assert sorted(m["VAR"] for m in tuning["0"]["candidates"]) == ["k-0-1x1x60"]
    assert [m["VAR"] for m in tuning["1"]["candidates"]] == ["k-1-1x1x60"]

    assert R.main([str(tmp_path)]) == 0
    out = capsys.readouterr().out
    assert '"spearman": 1.0' in out and "auto tuning" in out
    summary = R.read_rows(tmp_path / "summary.csv")
    assert len(summary) == 7 and float(summary[0]["NumMeasures"]) == 48
    svg = open(tmp_path / "scatter.svg").read()
    assert svg.startswith("<svg") and svg.count("<circle") == 2 and svg.count("<rect") == 2 + 5
    for label in ("opt", "hand", "auto", "min", "max"):
        assert ">%s<" % label in svg
