"""Multi-rank halo exchange: the protocol on CPU (gloo, world_size 2) and the full path on >= 2 GPUs."""

import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from absinthe_b200.distributed import SlabExchange, face_rows, neighbours

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_neighbours_wrap():
    assert neighbours(0, 4) == (3, 1) and neighbours(3, 4) == (2, 0) and neighbours(0, 1) == (0, 0)
    assert face_rows(64, 3) == ((3, 3), (64, 3), (0, 3), (67, 3))


class HostExchange(SlabExchange):
    """Test double: same protocol, fields are CPU tensors and the face kernels are slicing."""

    def __init__(self, fields, dims, world, rank):
        self.fields, self.dims, self.world, self.rank, self.pg, self.v = fields, dims, world, rank, None, None
        self.prev, self.next = neighbours(rank, world)
        self.buffers = {}

    def field_names(self, group):
        return sorted(self.fields)

    def allocate(self, count):
        X, Y, Z, HX, HY, HZ = self.dims
        return [torch.empty((count, Z + 2 * HZ, HY, X + 2 * HX), dtype=torch.float64) for _ in range(4)]

    def local_faces(self, name):
        X, Y, Z, HX, HY, HZ = self.dims
        f = self.fields[name]
        f[:, :, :HX] = f[:, :, X:X + HX]
        f[:, :, X + HX:] = f[:, :, HX:2 * HX]
        f[:HZ] = f[Z:Z + HZ]
        f[Z + HZ:] = f[HZ:2 * HZ]

    def pack(self, name, buffer, j0, width):
        buffer.copy_(self.fields[name][:, j0:j0 + width, :])

    def unpack(self, name, buffer, j0, width):
        self.fields[name][:, j0:j0 + width, :] = buffer


def _worker(rank, world, port, dims, failures):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, Y, Z, HX, HY, HZ = dims
    rng = np.random.default_rng(5)
    names = ["a", "b"]
    glob = {n: rng.standard_normal((Z, world * Y, X)) for n in names}
    fields = {}
    for n in names:
        local = np.zeros((Z + 2 * HZ, Y + 2 * HY, X + 2 * HX))
        local[HZ:HZ + Z, HY:HY + Y, HX:HX + X] = glob[n][:, rank * Y:(rank + 1) * Y, :]
        fields[n] = torch.from_numpy(local)
    HostExchange(fields, dims, world, rank).update(1)
    for n in names:
        padded = np.pad(glob[n], ((HZ, HZ), (HY, HY), (HX, HX)), mode="wrap")
        want = padded[:, rank * Y:rank * Y + Y + 2 * HY, :]
        if not np.array_equal(fields[n].numpy(), want):
            failures.put((rank, n))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_exchange_protocol_gloo(world):
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = 29500 + os.getpid() % 1000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, (12, 7, 5, 3, 3, 3), failures)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert failures.empty()


def test_single_rank_exchange_is_make_periodic():
    dims = (10, 6, 4, 3, 3, 3)
    X, Y, Z, HX, HY, HZ = dims
    rng = np.random.default_rng(1)
    interior = rng.standard_normal((Z, Y, X))
    local = np.zeros((Z + 6, Y + 6, X + 6))
    local[3:3 + Z, 3:3 + Y, 3:3 + X] = interior
    fields = {"f": torch.from_numpy(local)}
    HostExchange(fields, dims, 1, 0).update(1)
    assert np.array_equal(fields["f"].numpy(), np.pad(interior, 3, mode="wrap"))


@pytest.mark.gpu
def test_two_gpus_match_one(launcher_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "multi-gpu parity ok" in proc.stdout


@pytest.mark.gpu
def test_bench_variants_match_the_oracle_on_one_gpu(launcher_lib):
    """``bench.py --check`` at N = 1: the tuned 1024 x 1024 x 80 variants against the windowed oracle."""
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--check"], capture_output=True,
                          text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]
    report = json.loads(proc.stdout.strip().splitlines()[-1])
    assert report["check"] == "ok" and report["rows_compared"] > 0 and not report["failed"]


@pytest.mark.gpu
def test_bench_variants_match_the_oracle_across_two_slabs(launcher_lib):
    """The combination the scaling bench times -- tuned variants, fused periodic copies, per-group tile shapes,
    ``SlabExchange.update_async`` at 1024 x 1024 x 80 per rank -- with seam windows (the rows either side of both
    slab edges, halo rows included) compared with the oracle on every rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "bench.py"), "--check", "--gpus", "2"]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]
    report = json.loads([line for line in proc.stdout.splitlines() if line.startswith("{")][-1])
    assert report["check"] == "ok" and report["n_gpus"] == 2 and not report["failed"]


# ---- the exchange inside the launcher: P2P windows (absinthe_comm_*, put_group / wait_group) --------------
def _global_oracle(kernel, slab, world):
    from oracle.reference_oracle import Oracle
    from absinthe_b200 import programs as P
    from seam_check import field_seed, global_field
    X, Y, Z = slab
    extent = (X, world * Y, Z)
    oracle = Oracle(P.make_program(kernel, domain=extent))
    box = ((-3, X + 3), (-3, world * Y + 3), (-3, Z + 3))
    inputs = [global_field(field_seed(kernel, n), box, extent, "cpu").numpy() for n in oracle.inputs]
    return dict(zip(oracle.outputs, oracle.run(inputs)))


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["kernel", "copy", "fused"])
@pytest.mark.parametrize("kernel", ["diffusion", "fastwaves"])
def test_p2p_exchange_on_one_rank_is_the_periodic_update(launcher_lib, kernel, engine, tmp_path):
    from p2p_worker import worker
    slab = (48, 20, 10)
    worker(0, 1, 29621, kernel, slab, 2, str(tmp_path / "r%d.npz"), engine)
    want = _global_oracle(kernel, slab, 1)
    got = np.load(str(tmp_path / "r0.npz"))
    for name, w in want.items():
        assert np.array_equal(got[name], w), name


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,world,engine", [("diffusion", 2, None), ("fastwaves", 2, None), ("fastwaves", 3, None),
                                                 ("fastwaves", 2, "kernel"), ("diffusion", 3, "copy"), ("fastwaves", 3, "fused")])
def test_p2p_exchange_between_processes(launcher_lib, kernel, world, engine, tmp_path):
    """Ranks are separate processes (CUDA IPC windows); they may share one GPU, so this runs on a 1-GPU box too."""
    from p2p_worker import worker
    slab = (48, 20, 10)
    result = str(tmp_path / "r%d.npz")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=worker, args=(r, world, 29622 + world, kernel, slab, 3, result, engine))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    want = _global_oracle(kernel, slab, world)
    Y = slab[1]
    for r in range(world):
        got = np.load(result % r)
        for name, w in want.items():
            assert np.array_equal(got[name], w[:, r * Y:r * Y + Y + 6, :]), (name, r)


@pytest.mark.gpu
def test_second_put_without_wait_is_refused(launcher_lib):
    from absinthe_b200 import programs as P
    from absinthe_b200.distributed import P2PExchange
    from absinthe_b200.launcher import LauncherError, Variant
    v = Variant.from_program(P.make_program("advection", None, [(2, 2, 2)], domain=(40, 16, 8)))
    v.bind([v.empty_field() for _ in v.inputs], [v.empty_field() for _ in v.outputs])
    with pytest.raises(LauncherError, match="not connected"):
        v.put_group(1)
    exchange = P2PExchange(v, 1, 0)
    v.apply_group(1, with_halo=False)
    exchange.update_async(1)
    with pytest.raises(LauncherError, match="has not been waited for"):
        v.put_group(1)
    exchange.wait(1)
    exchange.wait(1)                                   # nothing outstanding: no-op
    exchange.update(1)
    torch.cuda.synchronize()
    v.close()


def test_p2p_exchange_picks_engines_and_agrees_on_failure(monkeypatch):
    """Host logic of P2PExchange without a GPU: the successor of a group reads its halos at once (fused push), the
    last group's exchange overlaps whatever comes next (copy engines); a connect failure raises a RuntimeError."""
    from absinthe_b200.distributed import P2PExchange

    class FakeVariant:
        n_groups, device = 3, 0
        calls = []

        def comm_create(self, rank, world):
            return b"h" * 64

        def comm_connect(self, prev, nxt):
            self.calls.append(("connect", prev, nxt))

        def apply_group_put(self, group, stream=None):
            self.calls.append(("fused", group))

        def apply_group(self, group, with_halo=True, stream=None):
            self.calls.append(("apply", group, with_halo))

    v = FakeVariant()
    ex = P2PExchange(v, 1, 0)
    assert ex.engines == {1: "kernel", 2: "kernel", 3: "copy"}
    ex.apply(1)
    assert v.calls[-1] == ("fused", 1)
    ex.update_async = lambda group, engine=None: v.calls.append(("put", group, engine))
    ex.apply(3)
    assert v.calls[-2:] == [("apply", 3, False), ("put", 3, "copy")]
    ex.apply(2, "kernel")
    assert v.calls[-2:] == [("apply", 2, False), ("put", 2, "kernel")]

    class Broken(FakeVariant):
        def comm_connect(self, prev, nxt):
            raise OSError("no peer access")
    with pytest.raises(RuntimeError, match="P2P exchange unavailable.*no peer access"):
        P2PExchange(Broken(), 1, 0)
