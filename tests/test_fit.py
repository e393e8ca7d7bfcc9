"""The LAD re-training of the model parameters (absinthe_b200/fit.py)."""

import os

import numpy as np
import pytest

from absinthe_b200 import fit


def synthetic_cache_rows(body, peel, cores, noise=0.0, seed=0):
    rng = np.random.default_rng(seed)
    rows = []
    for variant in ("PT8", "PT12", "PT16", "PT20"):
        fac = float(variant[2:])
        for xlen in (10, 20, 50):
            for ylen in (2, 5, 13):
                for zlen in (3, 8):
                    X, Y, Z = 2 * xlen, 2 * ylen, 5 * cores * zlen
                    t = cores * (body * fac * 9 * xlen * ylen * zlen + peel * fac * 9 * ylen * zlen)
                    rows.append((variant, X, Y, Z, t * (1 + noise * rng.standard_normal())))
    return rows


def test_lad_recovers_known_parameters():
    rows = synthetic_cache_rows(2e-8, 3e-7, cores=148)
    got = fit.fit_cache(rows, 148)
    assert got["BODY"] == pytest.approx(2e-8, rel=1e-6) and got["PEEL"] == pytest.approx(3e-7, rel=1e-6)
    assert got["R2"] > 0.999999
    assert got["SAMPLES"] == 54          # PT8 is excluded (fac >= 10), like the reference


def test_lad_is_robust_to_outliers():
    rows = synthetic_cache_rows(2e-8, 3e-7, cores=8, noise=0.01, seed=3)
    rows[5] = rows[5][:4] + (rows[5][4] * 50,)      # one wild sample
    got = fit.fit_cache(rows, 8)
    assert got["BODY"] == pytest.approx(2e-8, rel=0.05)


def test_ddr_design_matches_reference_formula():
    cores = 4
    rows = []
    true = {"RW BODY": 1e-7, "ST BODY": 4e-7, "RW PEEL": 2e-6, "ST PEEL": 5e-6}
    for variant in ("IN1BD0", "IN2BD1", "IN3BD2", "IN1BD2", "IN2BD0", "IN0BD0"):
        inp, bd = float(variant[2]), 2 * float(variant[5])
        for xlen, ylen, zlen in ((10, 5, 13), (20, 8, 5), (50, 3, 8), (80, 2, 5), (30, 13, 2)):
            xyz, xy, xz, yz = xlen * ylen * zlen, xlen * ylen, xlen * zlen, ylen * zlen
            t = cores * 9 * (true["RW BODY"] * xyz + true["ST BODY"] * ((inp + 1) * xyz + inp * bd * (xy + xz + yz))
                             + true["RW PEEL"] * yz + true["ST PEEL"] * ((inp + 1) * yz + inp * bd * (ylen + zlen)))
            rows.append((variant, 2 * xlen, 2 * ylen, 5 * cores * zlen, t))
    got = fit.fit_ddr(rows, cores)
    for key, value in true.items():
        assert got[key] == pytest.approx(value, rel=1e-5)
    assert got["SAMPLES"] == 25           # IN0BD0 is not trained on (ref fitddr.R:40)
    # a sweep that ran all 20 tiles per core (--stride 1) measures 20 x the per-tile time
    full = [r[:4] + (20 * r[4],) for r in rows]
    again = fit.fit_ddr(full, cores, tiles_per_core=20)
    for key, value in true.items():
        assert again[key] == pytest.approx(value, rel=1e-5)


def test_spearman():
    assert fit.spearman([1, 2, 3, 4], [10, 20, 30, 40]) == pytest.approx(1.0)
    assert fit.spearman([1, 2, 3, 4], [4, 3, 2, 1]) == pytest.approx(-1.0)
    assert fit.spearman([1, 2, 2, 3], [1, 2, 2, 3]) == pytest.approx(1.0)


@pytest.mark.skipif(not os.path.isfile("/root/reference/fitcache/results.csv"), reason="reference data not mounted")
def test_reproduces_readme_cache_parameters():
    # ref README.md:59-62
    got = fit.fit_cache(fit.load_results("/root/reference/fitcache/results.csv"), 4)
    assert got["BODY"] == pytest.approx(9.44349025644442e-08, rel=1e-5)
    assert got["PEEL"] == pytest.approx(9.95340542431222e-07, rel=1e-5)
