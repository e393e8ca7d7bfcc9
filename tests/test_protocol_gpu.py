"""The measurement protocol and the awkward corners of the variant space, on the GPU."""

import io

import numpy as np
import pytest

from absinthe_b200 import generator as G
from absinthe_b200 import programs as P
from common import assert_bit_exact, seq

pytestmark = pytest.mark.gpu


def test_runner_prints_the_reference_protocol(launcher_lib, tmp_path):
    from absinthe_b200.runner import run_variant
    import subprocess
    program = P.make_program("advection", [seq("advection")], [(1, 4, 5)], variant="advection-0-1-4-5", RUNS=5)
    stem = str(tmp_path / "advection-0-1-4-5")
    G.generate_code("template_tiling.cu", stem + ".cu", program)
    subprocess.check_call(G.nvcc_command(program["CUDA"]["NVCC_FLAGS"], stem + ".cu", stem + ".cubin"))
    out = io.StringIO()
    total, halo = run_variant(stem, 5, flush=True, out=out)
    assert len(total) == 5 and np.all(total > 0) and np.all(halo > 0) and np.all(total > halo)   # TOTAL includes HALO
    log = tmp_path / "output.txt"
    log.write_text(out.getvalue())
    rows = G.parse_results(str(log), 0)
    assert len(rows) == 5 and rows[0][:4] == ["advection-0-1-4-5", "64", "64", "60"]
    assert len(G.parse_results(str(log), 2)) == 3


CORNERS = [
    # more tiles than points along x and z: most tiles are empty after clipping
    ("more-tiles-than-points", P.make_program("advection", [seq("advection")], [(24, 2, 9)], domain=(12, 10, 4))),
    # extents smaller than the halo: the literal three-pass periodic update is needed
    ("extent-below-halo", P.make_program("diffusion", [seq("diffusion")], [(1, 1, 1)], domain=(8, 2, 2))),
    ("single-point-tiles", P.make_program("fastwaves", [seq("fastwaves")[:6], seq("fastwaves")[6:]],
                                          [(8, 6, 4), (1, 1, 1)], domain=(8, 6, 4))),
    ("one-long-row", P.make_program("advection", [seq("advection")], [(1, 1, 1)], domain=(600, 4, 3))),
]


@pytest.mark.parametrize("name,program", CORNERS, ids=[c[0] for c in CORNERS])
@pytest.mark.parametrize("cuda", [{}, {"INLINE": "all", "BLOCK": (2, 2), "SHUFFLE": True, "FUSE_HALO": True}],
                         ids=["generic", "tuned-options"])
def test_corner_cases_match_oracle(launcher_lib, name, program, cuda):
    program = P.clone(program)
    program["CUDA"] = dict(cuda)
    assert_bit_exact(program)


def test_unbound_variant_is_an_error(launcher_lib):
    from absinthe_b200.launcher import LauncherError, Variant
    v = Variant.from_program(P.make_program("advection", domain=(16, 8, 4)))
    with pytest.raises(LauncherError, match="no buffers bound"):
        v.apply()
    v.close()


@pytest.mark.parametrize("kernel,outputs", [("diffusion", 4), ("advection", 2), ("fastwaves", 3)])
def test_verify_runs_the_sequential_reference(launcher_lib, tmp_path, kernel, outputs):
    """VERIFY: the variant is compared with the per-stencil sequential variant (ref template_tiling.cpp:378-396)."""
    from absinthe_b200.runner import run_variant
    import subprocess
    program = P.make_program(kernel, [seq(kernel)], [(2, 4, 6)], variant=kernel + "-fused", RUNS=2, VERIFY=True)
    program["CUDA"] = {"INLINE": "all", "BLOCK": (2, 1), "SHUFFLE": True, "FUSE_HALO": True}
    stem = str(tmp_path / kernel)
    G.generate_code("template_tiling.cu", stem + ".cu", program)
    for s in (stem, stem + G.REFERENCE_SUFFIX):
        subprocess.check_call(G.nvcc_command(program["CUDA"]["NVCC_FLAGS"], s + ".cu", s + ".cubin"))
    out = io.StringIO()
    run_variant(stem, 2, out=out, verify=True)
    lines = out.getvalue().splitlines()
    assert "   - verify True" in lines and "-> computing reference..." in lines
    assert lines[-3:] == ["-> verifying...", "   - 0 errors", "   - %d matches" % (outputs * 64 * 64 * 60)]


def test_compare_fields_counts_like_the_reference(launcher_lib):
    import torch
    from absinthe_b200.launcher import compare_fields
    dims = (20, 6, 5, 3, 3, 3)
    shape = (5 + 6, 6 + 6, 20 + 6)
    gen = torch.Generator(device="cuda").manual_seed(5)
    a = torch.rand(shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    b = a.clone()
    assert compare_fields(a, b, dims) == (0, 20 * 6 * 5)
    b[3 + 2, 3 + 1, 3 + 7] *= 1.0 + 3e-7          # beyond the 1e-7 relative tolerance
    b[3 + 0, 3 + 0, 3 + 0] *= 1.0 + 5e-8          # within it
    b[0, 0, 0] = 99.0                             # halo: not looked at
    assert compare_fields(a, b, dims) == (1, 20 * 6 * 5 - 1)
    assert compare_fields(a, b, dims, tolerance=1e-8) == (2, 20 * 6 * 5 - 2)
    # the reference's test never fires on NaN (ref template_tiling.cpp:386), so an all-NaN output would verify as
    # "0 errors"; here a non-finite value where the expected one is finite is an error
    b[3 + 4, 3 + 5, 3 + 19] = float("nan")
    b[3 + 4, 3 + 5, 3 + 18] = float("inf")
    assert compare_fields(a, b, dims) == (3, 20 * 6 * 5 - 3)
    a[3 + 4, 3 + 5, 3 + 19] = float("nan")        # expected NaN as well: not this check's business
    assert compare_fields(a, b, dims) == (2, 20 * 6 * 5 - 2)


def test_runner_drives_two_ranks(launcher_lib, tmp_path):
    """``run.sh`` of a program with TILING.NY = 2: torchrun, one process per rank (they may share a GPU), P2P halo
    exchange inside the launcher; rank 0 prints the reference's protocol and parse_results reads it."""
    import subprocess
    import sys
    from absinthe_b200 import generator as G
    program = P.make_program("fastwaves", variant="fw-2ranks", domain=(64, 48, 12), RUNS=3)
    program["TILING"]["NY"] = 2
    program["CUDA"] = {"INLINE": "all", "FUSE_HALO": True}
    stem = str(tmp_path / "fw-2ranks")
    flags = P.clone(program)
    G.generate_code("template_tiling.cu", stem + ".cu", flags)
    subprocess.check_call(G.nvcc_command(flags["CUDA"]["NVCC_FLAGS"], stem + ".cu", stem + ".cubin"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", "-m", "absinthe_b200.runner", stem]
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root,
                          env=dict(os.environ, PYTHONPATH=root))
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]
    log = tmp_path / "log.txt"
    log.write_text(proc.stdout)
    rows = G.parse_results(str(log), 0)
    assert len(rows) == 3 and rows[0][0] == "fw-2ranks" and rows[0][1:4] == ["64", "48", "12"]
    assert all(float(r[4]) >= float(r[5]) >= 0 for r in rows)
