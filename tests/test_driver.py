"""The experiment drivers keep the reference's command-line recipe (-a/-g/-b/-p/-f)."""

import os
import subprocess
import sys

import pytest

from absinthe_b200 import driver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_auto_generate_writes_variants_makefile_and_script(tmp_path):
    folder = str(tmp_path / "advection")
    experiments = driver.main("advection", ["-a", "-g", "-f", folder])
    # ref advection.py:102-107: nx, ny in 1..64, nz in {1,2,5,15,30,60}, at least CORES (= 148 SMs) tiles
    assert len(experiments) == 198 and all(n.startswith("advection-0-") for n in experiments)
    for name in list(experiments)[:3]:
        assert os.path.exists(os.path.join(folder, name + ".cu")) and os.path.exists(os.path.join(folder, name + ".desc"))
    make = open(os.path.join(folder, "Makefile")).read()
    assert "advection-0-64-64-60.cubin: advection-0-64-64-60.cu" in make
    # the Makefile rule really builds a cubin for sm_100a
    target = "advection-0-4-8-5.cubin"
    subprocess.check_call(["make", target], cwd=folder, stdout=subprocess.DEVNULL)
    assert os.path.getsize(os.path.join(folder, target)) > 1000
    script = open(os.path.join(folder, "run.sh")).read()
    assert "absinthe_b200.runner" in script and "./advection-0-4-8-5" in script and "PYTHONPATH=" + ROOT in script


def test_fastwaves_auto_is_unfiltered_like_the_reference():
    experiments = {}
    driver.auto_tune("fastwaves", experiments)
    assert len(experiments) == 2 * 7 * 7 * 6          # ref fastwaves.py:100: the >= CORES filter is commented out
    assert experiments["fastwaves-1-1-1-1"]["OUTPUTS"] == ["div"]


def test_parse_flag_writes_auto_csv(tmp_path):
    folder = tmp_path / "diffusion"
    folder.mkdir()
    log = open(os.path.join(ROOT, "tests", "golden", "sample_output.txt")).read()
    (folder / "output.txt").write_text(log)
    driver.main("diffusion", ["-a", "-p", "output.txt", "-f", str(folder)])
    rows = (folder / "auto.csv").read_text().splitlines()
    assert rows[0] == "VAR,X,Y,Z,TOTAL,HALO" and len(rows) == 6
    driver.main("diffusion", ["-p", "output.txt", "-f", str(folder)])          # default skip = 16 -> no rows
    assert (folder / "results.csv").read_text().splitlines() == ["VAR,X,Y,Z,TOTAL,HALO"]


def test_training_drivers_enumerate_the_reference_sweep():
    cache = driver.training_experiments("fitcache", cores=4)
    ddr = driver.training_experiments("fitddr", cores=4)
    assert len(cache) == 4 * 103 and len(ddr) == 10 * 103       # ref fitcache.py:116-172, fitddr.py:187-327
    p = cache["PT12-100x10x60"]          # the first sample of the reference's fitcache/results.csv
    assert (p["X"], p["Y"], p["Z"]) == (100, 10, 60) and p["TILING"]["GROUPS"][0]["GROUPS"][0]["NZ"] == 20


def test_training_stride_option(tmp_path):
    """-s/--stride: every N-th tile is executed (20 = the reference's one-tile-per-core protocol)."""
    from absinthe_b200 import generator as G
    default = driver.training_experiments("fitcache", cores=4)["PT12-100x10x60"]
    assert default["TRAINING"]["STRIDE"] == 20
    full = driver.training_experiments("fitcache", cores=4, stride=1)["PT12-100x10x60"]
    assert full["TRAINING"]["STRIDE"] == 1
    G.generate_code("template_training.cu", str(tmp_path / "t.cu"), full)
    assert " training_step 1 " in (tmp_path / "t.desc").read_text()


def test_root_scripts_run(tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "fitcache.py"), "-f", str(tmp_path)],
                         capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "-> working dir" in out.stdout
