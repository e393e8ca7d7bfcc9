"""Launcher robustness: load failures are errors (not crashes), the async host path refuses
pageable memory, per-device scratch, variants on a device other than the current one."""

import ctypes

import pytest
import torch

from absinthe_b200 import programs as P
from absinthe_b200.generator import build_variant
from absinthe_b200.launcher import LauncherError, Variant, compare_fields, flush_l2, load_library, make_periodic

pytestmark = pytest.mark.gpu


def small():
    return P.make_program("advection", None, [(2, 4, 3)], domain=(40, 24, 10))


def test_load_failures_are_reported_not_fatal(launcher_lib, tmp_path):
    desc, cubin = build_variant(small())
    text = open(desc).read()
    bad = tmp_path / "bad.desc"
    bad.write_text(text.replace("kernel absinthe_advection_g1", "kernel no_such_kernel"))
    with pytest.raises(LauncherError, match="no_such_kernel"):
        Variant(str(bad), cubin)
    greedy = tmp_path / "greedy.desc"
    import re
    greedy.write_text(re.sub(r"smem \d+", "smem 400000", text))
    with pytest.raises(LauncherError, match="shared memory|resident"):
        Variant(str(greedy), cubin)
    garbage = tmp_path / "garbage.cubin"
    garbage.write_bytes(b"not a cubin" * 100)
    with pytest.raises(LauncherError, match="cuModuleLoadData"):
        Variant(desc, str(garbage))
    Variant(desc, cubin).close()                      # the library is still usable afterwards


def test_async_host_path_refuses_pageable_memory(launcher_lib):
    v = Variant.from_program(small())
    v.bind([v.empty_field() for _ in v.inputs], [v.empty_field() for _ in v.outputs])
    pinned_in = [torch.zeros(v.shape, dtype=torch.float64).pin_memory() for _ in v.inputs]
    pinned_out = [torch.zeros(v.shape, dtype=torch.float64).pin_memory() for _ in v.outputs]
    pageable = [torch.zeros(v.shape, dtype=torch.float64) for _ in v.outputs]
    with pytest.raises(LauncherError, match="page-locked"):
        v.apply_host(pinned_in, pageable, wait=False)
    with pytest.raises(LauncherError, match="page-locked"):
        v.apply_host([torch.zeros(v.shape, dtype=torch.float64) for _ in v.inputs], pinned_out, wait=False)
    v.apply_host(pinned_in, pageable)                 # the synchronous call takes any host memory
    v.apply_host(pinned_in, pinned_out, wait=False)
    v.wait_host()
    v.close()


def test_flush_and_compare_reuse_their_scratch(launcher_lib):
    a = torch.rand(16, 30, 46, dtype=torch.float64, device="cuda")
    dims = (40, 24, 10, 3, 3, 3)
    make_periodic(a, dims)
    b = a.clone()
    b[5, 7, 9] += 1.0
    for _ in range(3):                                # the counters are allocated once and zeroed per call
        assert compare_fields(a, a, dims) == (0, 40 * 24 * 10)
        assert compare_fields(a, b, dims) == (1, 40 * 24 * 10 - 1)
        flush_l2()
    torch.cuda.synchronize()
    free_before = torch.cuda.mem_get_info()[0]
    for _ in range(5):
        flush_l2()
        compare_fields(a, b, dims)
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] == free_before


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_variant_on_another_device_than_the_current_one(launcher_lib):
    from oracle.reference_oracle import Oracle
    program = small()
    oracle = Oracle(P.clone(program))
    inputs = oracle.fill_inputs(0)
    expected = oracle.run(inputs)
    v = Variant.from_program(P.clone(program), device=1)
    torch.cuda.set_device(0)                          # the caller's current device is not the variant's
    by_name = dict(zip(oracle.inputs, inputs))
    ins = [torch.from_numpy(by_name[n]).to("cuda:1") for n in v.inputs]
    outs = [v.empty_field() for _ in v.outputs]
    v.bind(ins, outs)
    v.apply()
    total, _ = v.run(2, flush=True)
    torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0 and len(total) == 2
    for name, want in zip(oracle.outputs, expected):
        assert outs[v.outputs.index(name)].cpu().numpy().tobytes() == want.tobytes()
    v.close()


def test_cluster_kernels_launch_whole_clusters_and_refuse_the_training_loop(launcher_lib, tmp_path):
    """COLUMN.CLUSTER: the grid is a whole number of co-resident clusters (graph replay included), results are the
    oracle's, and the training loop over single tiles is refused; a cluster beyond the portable size is a load error."""
    import re
    from common import assert_bit_exact
    f = P.KERNELS["fastwaves"][1]()
    program = P.make_program("fastwaves", [f], [(2, 4, 6)], domain=(70, 26, 12))     # 3 x 13 tiles: dead ranks in x and y
    program["CUDA"] = {"STREAM_K": "regs", "FUSE_HALO": True,
                       "COLUMN": {"MERGE": True, "WARPS": (1, 2), "UNROLL": 1, "CLUSTER": (2, 2), "SLACK": 1}}
    assert_bit_exact(program)
    assert_bit_exact(program, graph=True)
    v = Variant.from_program(P.clone(program))
    v.bind([v.empty_field() for _ in v.inputs], [v.empty_field() for _ in v.outputs])
    with pytest.raises(LauncherError, match="training loop"):
        v.run(1, flush=False, training=True)
    v.run(2, flush=False)
    v.close()
    desc, cubin = build_variant(P.clone(program))
    big = tmp_path / "big.desc"
    big.write_text(re.sub(r"cluster \d+", "cluster 16", open(desc).read()))
    with pytest.raises(LauncherError, match="portable limit"):
        Variant(str(big), cubin)
