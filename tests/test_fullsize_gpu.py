"""Parity at BASELINE.json's full size (1024 x 1024 x 80 per GPU) through size-independent properties:
the oracle cannot run there in seconds, so the tuned bench variants are checked by
  * a windowed oracle  -- the oracle on small windows cut out of the full-size inputs reproduces the GPU
                          result wherever the window's own periodic wrap cannot reach,
  * variant invariance -- a differently fused / tiled / un-inlined variant gives the same bits,
  * translation        -- rolling the periodic inputs rolls the outputs,
  * periodicity        -- every output halo equals the wrapped interior,
and the end-to-end host path returns the same bits as the device-resident path."""

import numpy as np
import pytest
import torch

from absinthe_b200 import programs as P
from absinthe_b200.launcher import Variant, make_periodic

pytestmark = pytest.mark.gpu

DOMAIN = (1024, 1024, 80)
MARGIN = 6          # > total dependency radius of either sequence (<= 3 per group, two groups at most)


def tuned_programs():
    import bench
    return dict(bench.workload_programs(DOMAIN))


def alternative(kernel):
    seq = P.KERNELS[kernel][1]()
    if kernel == "diffusion":      # per-field groups, intermediates in shared memory, separate halo kernels
        return P.make_program(kernel, P.split_sequence(seq, [0] * 4 + [1] * 4 + [2] * 4 + [3] * 4),
                              [(8, 64, 80)] * 4, domain=DOMAIN)
    return P.make_program(kernel, [seq], [(16, 32, 5)], domain=DOMAIN)          # fully fused fast waves


def run(program, inputs):
    v = Variant.from_program(P.clone(program))
    outs = [v.empty_field() for _ in v.outputs]
    v.bind([inputs[n] for n in v.inputs], outs)
    v.apply()
    torch.cuda.synchronize()
    names = list(v.outputs)
    v.close()
    return dict(zip(names, outs))


def make_inputs(program, seed):
    v = Variant.from_program(P.clone(program))
    dims = (v.X, v.Y, v.Z, v.HX, v.HY, v.HZ)
    gen = torch.Generator(device="cuda").manual_seed(seed)
    fields = {}
    for name in v.inputs:
        t = torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) + 0.5
        make_periodic(t, dims)
        fields[name] = t
    v.close()
    return fields, dims


@pytest.mark.parametrize("kernel", ["diffusion", "fastwaves"])
def test_full_size_properties(launcher_lib, kernel):
    from oracle.reference_oracle import Oracle
    program = tuned_programs()[kernel]
    inputs, dims = make_inputs(program, seed=11)
    X, Y, Z, H = dims[0], dims[1], dims[2], dims[3]
    got = run(program, inputs)

    # periodicity of every output: halos are wrapped copies of the interior
    for name, t in got.items():
        again = t.clone()
        make_periodic(again, dims)
        assert torch.equal(again, t), "%s: halo is not the wrapped interior" % name

    # variant invariance
    other = run(alternative(kernel), inputs)
    for name, t in got.items():
        assert torch.equal(other[name], t), "%s differs between variants" % name
    del other

    # windowed oracle at a corner, an edge and the middle of the domain
    wx, wy, wz = 40, 28, 14
    for ox, oy, oz in ((0, 0, 0), (X - wx, 500, Z - wz), (491, 333, 29)):
        small = P.make_program(kernel, domain=(wx, wy, wz))
        oracle = Oracle(small)
        # padded window = interior window + its true halo from the full-size (periodic) field
        win = [inputs[n][oz:oz + wz + 2 * H, oy:oy + wy + 2 * H, ox:ox + wx + 2 * H].cpu().numpy().copy()
               for n in oracle.inputs]
        want = oracle.run(win)
        m = MARGIN
        for name, w in zip(oracle.outputs, want):
            g = got[name][oz:oz + wz + 2 * H, oy:oy + wy + 2 * H, ox:ox + wx + 2 * H].cpu().numpy()
            core = (slice(H + m, H + wz - m), slice(H + m, H + wy - m), slice(H + m, H + wx - m))
            assert np.array_equal(g[core], w[core]), "%s window at %s differs from the oracle" % (name, (ox, oy, oz))

    # translation: roll the periodic interior, outputs roll with it
    shift = (5, -17, 130)          # (k, j, i)
    rolled = {}
    for name, t in inputs.items():
        r = t.clone()
        interior = t[H:H + Z, H:H + Y, H:H + X]
        r[H:H + Z, H:H + Y, H:H + X] = torch.roll(interior, shift, dims=(0, 1, 2))
        make_periodic(r, dims)
        rolled[name] = r
    moved = run(program, rolled)
    for name, t in got.items():
        want = torch.roll(t[H:H + Z, H:H + Y, H:H + X], shift, dims=(0, 1, 2))
        assert torch.equal(moved[name][H:H + Z, H:H + Y, H:H + X], want), "%s is not translation equivariant" % name


def test_host_path_matches_device_path(launcher_lib):
    program = P.make_program("fastwaves", [P.KERNELS["fastwaves"][1]()[:6], P.KERNELS["fastwaves"][1]()[6:]],
                             [(2, 8, 4), (4, 4, 2)], domain=(128, 96, 20))
    inputs, _ = make_inputs(program, seed=3)
    want = run(program, inputs)
    v = Variant.from_program(P.clone(program))
    dev_in = [torch.zeros_like(inputs[n]) for n in v.inputs]
    dev_out = [v.empty_field() for _ in v.outputs]
    v.bind(dev_in, dev_out)
    host_in = [inputs[n].cpu().pin_memory() for n in v.inputs]
    host_out = [torch.empty(v.shape, dtype=torch.float64).pin_memory() for _ in v.outputs]
    v.apply_host(host_in, host_out)
    for name, t in zip(v.outputs, host_out):
        assert torch.equal(t, want[name].cpu()), name
    for t in host_out:
        t.zero_()
    v.apply_host(host_in, host_out, wait=False)
    v.wait_host()
    for name, t in zip(v.outputs, host_out):
        assert torch.equal(t, want[name].cpu()), name
    # constants stay resident: a None entry skips the upload and uses the field on the device
    for t in host_out:
        t.zero_()
    resident = ["wgtfac", "hhl", "rho", "dzdx", "dzdy"]
    v.apply_host([None if n in resident else h for n, h in zip(v.inputs, host_in)], host_out)
    for name, t in zip(v.outputs, host_out):
        assert torch.equal(t, want[name].cpu()), name
    v.close()


def test_two_host_steps_in_flight(launcher_lib):
    """The async host path keeps two steps in flight; each step's outputs match the synchronous call's."""
    from absinthe_b200.launcher import LauncherError
    program = P.make_program("diffusion", None, [(2, 8, 4)], domain=(96, 64, 20))
    v = Variant.from_program(P.clone(program))
    v.bind([v.empty_field() for _ in v.inputs], [v.empty_field() for _ in v.outputs])
    steps = []
    for seed in (1, 2, 3, 4):
        inputs, _ = make_inputs(program, seed=seed)
        host_in = [inputs[n].cpu().pin_memory() for n in v.inputs]
        want = [torch.empty(v.shape, dtype=torch.float64).pin_memory() for _ in v.outputs]
        v.apply_host(host_in, want)
        steps.append((host_in, want))
    outs = [[torch.zeros(v.shape, dtype=torch.float64).pin_memory() for _ in v.outputs] for _ in range(2)]
    v.apply_host(steps[0][0], outs[0], wait=False)
    v.apply_host(steps[1][0], outs[1], wait=False)
    with pytest.raises(LauncherError):
        v.apply_host(steps[2][0], outs[0], wait=False)       # a third step needs a wait first
    for n in range(4):
        v.wait_host()
        for name, got, want in zip(v.outputs, outs[n % 2], steps[n][1]):
            assert torch.equal(got, want), (n, name)
        if n + 2 < 4:
            v.apply_host(steps[n + 2][0], outs[n % 2], wait=False)
    v.wait_host()                                            # nothing pending: returns at once
    # a synchronous call drains steps still in flight before it touches the device fields
    v.apply_host(steps[0][0], outs[0], wait=False)
    v.apply_host(steps[1][0], outs[1])
    for got, want in zip(outs[0], steps[0][1]):
        assert torch.equal(got, want)
    for got, want in zip(outs[1], steps[1][1]):
        assert torch.equal(got, want)
    v.close()


def test_interior_upload_rebuilds_the_halos(launcher_lib):
    """host_interior: only interiors cross the bus, the periodic halos of the inputs are rebuilt on the device."""
    program = P.make_program("advection", None, [(2, 4, 3)], domain=(70, 33, 11))
    inputs, _ = make_inputs(program, seed=9)
    want = run(program, inputs)
    v = Variant.from_program(P.clone(program))
    v.bind([v.empty_field() for _ in v.inputs], [v.empty_field() for _ in v.outputs])
    v.host_interior(True)
    host_in = []
    for name in v.inputs:
        h = inputs[name].cpu()
        spoiled = torch.full_like(h, 1e300)                  # whatever the caller left in the halos is ignored
        spoiled[3:-3, 3:-3, 3:-3] = h[3:-3, 3:-3, 3:-3]
        host_in.append(spoiled.pin_memory())
    for wait in (True, False):
        host_out = [torch.zeros(v.shape, dtype=torch.float64).pin_memory() for _ in v.outputs]
        v.apply_host(host_in, host_out, wait=wait)
        v.wait_host()
        for name, t in zip(v.outputs, host_out):
            assert torch.equal(t, want[name].cpu()), (wait, name)
    v.host_interior(False)                                   # back to whole fields: the spoiled halos now count
    host_out = [torch.zeros(v.shape, dtype=torch.float64).pin_memory() for _ in v.outputs]
    v.apply_host(host_in, host_out)
    assert any(not torch.equal(t, want[name].cpu()) for name, t in zip(v.outputs, host_out))
    v.close()
