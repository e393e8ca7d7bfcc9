"""Synthetic periodic input fields that every rank of a decomposed run can build on its own.

The reference fills its inputs with ``rand()`` in one process (ref template_tiling.cpp:62-64, 209-221).  With the
domain cut into j-slabs over several GPUs a rank needs its slab *and* the halo rows that belong to its neighbours;
an exact integer hash of the global periodic coordinates gives both without any communication (and lets a checker
evaluate any window of the global field).
"""


def hashed(gi, gj, gk, seed):
    """Exact pseudo-random doubles in [0.5, 1.5) from integer coordinates (torch int64 tensors, broadcast)."""
    v = (gi * 73856093) ^ (gj * 19349663) ^ (gk * 83492791) ^ (seed * 2654435761)
    v = (v ^ (v >> 13)) * 1274126177
    return ((v ^ (v >> 16)) & 0xFFFFF).double() / float(1 << 20) + 0.5


def global_field(seed, box, extent, device):
    """Padded field over the global index box ``(i0, i1), (j0, j1), (k0, k1)`` of a periodic ``extent`` domain."""
    import torch
    (i0, i1), (j0, j1), (k0, k1) = box
    gi = torch.arange(i0, i1, device=device, dtype=torch.int64) % extent[0]
    gj = torch.arange(j0, j1, device=device, dtype=torch.int64) % extent[1]
    gk = torch.arange(k0, k1, device=device, dtype=torch.int64) % extent[2]
    return hashed(gi[None, None, :], gj[None, :, None], gk[:, None, None], seed).contiguous()


def field_seed(kernel, name):
    return sum(ord(c) * (n + 1) for n, c in enumerate(kernel + ":" + name)) % 100003


def slab_field(kernel, name, slab, halo, rank, world, device):
    """Input ``name`` of rank ``rank``: its ``slab`` = (X, Y, Z) plus halos, cut out of the X x (world*Y) x Z field."""
    X, Y, Z = slab
    HX, HY, HZ = halo
    box = ((-HX, X + HX), (rank * Y - HY, rank * Y + Y + HY), (-HZ, Z + HZ))
    return global_field(field_seed(kernel, name), box, (X, world * Y, Z), device)
