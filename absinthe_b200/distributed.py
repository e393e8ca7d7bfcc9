"""Horizontal domain decomposition across the GPUs of one box: j-slabs with NVLink halo exchange.

The reference never filled in its distributed hooks -- ``TILING.NX/NY/NZ`` with
``index[] = {0,0,0}`` (ref template_tiling.cpp:189-203), per-group ``HALOS`` and the ``PUT``/``WAIT``
schedule slots (ref stencil_generator.py:23-37).  Here they mean: rank ``r`` of ``N`` owns the j-slab
``[r*Y, (r+1)*Y)`` of a global ``X x (N*Y) x Z`` periodic domain (k stays whole, i-rows stay
contiguous).  After a group's kernel, each of its outputs gets

1. the local periodic update of its i and k faces (``absinthe_make_periodic_ik``),
2. a PUT of its first / last ``HY`` interior rows -- every plane, full padded width, so the
   neighbours also receive consistent i/k halos -- to ranks ``r-1`` / ``r+1`` (mod N), and
3. a WAIT + unpack of the rows the neighbours put into its low / high j-halo.

At N = 1 both neighbours are the rank itself and the result is exactly ``make_periodic``.
Faces are packed by a kernel of the launcher into one contiguous buffer per direction (all outputs
of the group together) and moved with ``torch.distributed`` P2P (NCCL over NVLink on the GPUs,
gloo in the CPU tests of the protocol).
"""


def neighbours(rank, world):
    """(previous, next) rank along j with periodic wrap."""
    return (rank - 1) % world, (rank + 1) % world


def face_rows(Y, HY):
    """Padded row ranges: (low interior, high interior, low halo, high halo), each ``(j0, width)``."""
    return (HY, HY), (Y, HY), (0, HY), (Y + HY, HY)


class SlabExchange:
    """Halo exchange of one variant's group outputs between the j-neighbour ranks."""

    def __init__(self, variant, world, rank, process_group=None):
        self.v, self.world, self.rank, self.pg = variant, world, rank, process_group
        self.prev, self.next = neighbours(rank, world)
        self.dims = (variant.X, variant.Y, variant.Z, variant.HX, variant.HY, variant.HZ)
        self.buffers = {}

    # -- device plumbing (overridden by the CPU protocol test) ----------------------------
    def field_names(self, group):
        return self.v.group_outputs(group)

    def allocate(self, count):
        import torch
        X, Y, Z, HX, HY, HZ = self.dims
        shape = (count, Z + 2 * HZ, HY, X + 2 * HX)
        dev = "cuda:%d" % self.v.device
        return [torch.empty(shape, dtype=torch.float64, device=dev) for _ in range(4)]

    def local_faces(self, name):
        from .launcher import check, load_library, _stream_handle
        lib = load_library()
        check(lib, lib.absinthe_make_periodic_ik(self.v.field_ptr(name), *self.dims, _stream_handle(None)),
              "absinthe_make_periodic_ik")

    def pack(self, name, buffer, j0, width):
        from .launcher import check, load_library, _stream_handle
        lib = load_library()
        check(lib, lib.absinthe_halo_pack(self.v.field_ptr(name), buffer.data_ptr(), *self.dims, j0, width,
                                          _stream_handle(None)), "absinthe_halo_pack")

    def unpack(self, name, buffer, j0, width):
        from .launcher import check, load_library, _stream_handle
        lib = load_library()
        check(lib, lib.absinthe_halo_unpack(self.v.field_ptr(name), buffer.data_ptr(), *self.dims, j0, width,
                                            _stream_handle(None)), "absinthe_halo_unpack")

    def pack_all(self, group, names, send_low, send_high):
        """Local i/k faces + both j-faces of all outputs of the group (one launch each on the GPU)."""
        if hasattr(self.v, "pack_group"):
            self.v.pack_group(group, send_low, send_high, local_faces=True)
            return
        low_int, high_int, _, _ = face_rows(self.dims[1], self.dims[4])
        for n, name in enumerate(names):
            self.local_faces(name)
            self.pack(name, send_low[n], *low_int)
            self.pack(name, send_high[n], *high_int)

    def unpack_all(self, group, names, recv_low, recv_high):
        if hasattr(self.v, "unpack_group"):
            self.v.unpack_group(group, recv_low, recv_high)
            return
        _, _, low_halo, high_halo = face_rows(self.dims[1], self.dims[4])
        for n, name in enumerate(names):
            self.unpack(name, recv_low[n], *low_halo)
            self.unpack(name, recv_high[n], *high_halo)

    # -- overlap ------------------------------------------------------------------------------
    def update_async(self, group):
        """PUT on a side stream: the exchange of ``group`` runs concurrently with whatever the caller
        launches next on its own stream; ``wait(group)`` is the matching WAIT (ref schedule slots,
        stencil_generator.py:23-37)."""
        import torch
        if not hasattr(self, "comm"):
            self.comm = torch.cuda.Stream(device=self.v.device)
            self.done = {}
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            self.update(group)
            done = torch.cuda.Event()
            done.record(self.comm)
        self.done[group] = done

    def wait(self, group):
        """Make the caller's stream wait for the outstanding exchange of ``group`` (if any)."""
        import torch
        done = getattr(self, "done", {}).pop(group, None)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)

    def apply(self, group, engine=None):
        """COMP + PUT of ``group`` (same call as ``P2PExchange.apply``)."""
        self.v.apply_group(group, with_halo=False)
        self.update_async(group)

    # -- protocol ---------------------------------------------------------------------------
    def update(self, group):
        """PUT + WAIT of every output of ``group`` (the slot after its COMP in the schedule)."""
        import torch.distributed as dist
        names = self.field_names(group)
        if not names:
            return
        if group not in self.buffers:
            self.buffers[group] = self.allocate(len(names))
        send_low, send_high, recv_low, recv_high = self.buffers[group]
        self.pack_all(group, names, send_low, send_high)
        if self.world == 1:
            recv_high.copy_(send_low)
            recv_low.copy_(send_high)
        else:
            ops = [dist.P2POp(dist.isend, send_low, self.prev, self.pg),
                   dist.P2POp(dist.isend, send_high, self.next, self.pg),
                   dist.P2POp(dist.irecv, recv_high, self.next, self.pg),
                   dist.P2POp(dist.irecv, recv_low, self.prev, self.pg)]
            for work in dist.batch_isend_irecv(ops):
                work.wait()
        self.unpack_all(group, names, recv_low, recv_high)


class P2PExchange:
    """The same PUT / WAIT protocol inside the launcher: peer-to-peer stores over NVLink, no collective library.

    Every rank creates an exchange window on its GPU (``absinthe_comm_create``), the 64-byte IPC handles travel
    once through ``torch.distributed`` (any backend: they are host bytes), and from then on a PUT is one kernel
    that stores this rank's edge rows straight into the neighbours' windows and raises their flags, a WAIT one
    kernel that waits for the flags and fills the j-halos (``absinthe_variant_put_group`` / ``_wait_group``).
    Interface of ``SlabExchange``: ``update_async(group)`` = PUT on a side stream, ``wait(group)`` = WAIT on the
    caller's stream, ``update(group)`` = both.
    """

    def __init__(self, variant, world, rank, process_group=None):
        import torch.distributed as dist
        self.v, self.world, self.rank = variant, world, rank
        self.prev, self.next = neighbours(rank, world)
        mine = variant.comm_create(rank, world)
        handles = [mine]
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, mine, group=process_group)
        problem = None
        try:
            variant.comm_connect(handles[self.prev], handles[self.next])
        except Exception as exc:                       # e.g. no peer access between two devices
            problem = "rank %d: %s" % (rank, exc)
        if world > 1:
            # every rank learns whether all windows are mapped (this is also the barrier before the first PUT);
            # a failure anywhere raises everywhere, so that the caller can fall back on all ranks alike
            problems = [None] * world
            dist.all_gather_object(problems, problem, group=process_group)
            problem = next((p for p in problems if p), None)
        if problem:
            raise RuntimeError("P2P exchange unavailable (%s)" % problem)
        self.comm = None
        # default engine of update_async per group: the successor of a group reads its halos at once
        self.engines = {g: ("kernel" if g < variant.n_groups else "copy") for g in range(1, variant.n_groups + 1)}

    def update_async(self, group, engine=None):
        """PUT on a side stream.  ``engine``: "kernel" when the very next kernel reads these halos (lowest latency),
        "copy" when the exchange overlaps a compute kernel that does not (the copy engines take no SM resources
        from it).  Default: "kernel" for the last group of the variant and for every group but the last of a
        multi-group variant (its successor reads the halos), "copy" otherwise."""
        import torch
        if self.comm is None:
            self.comm = torch.cuda.Stream(device=self.v.device)
        if engine is None:
            engine = self.engines.get(group, "kernel")
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.v.device))
        self.comm.wait_event(ready)
        self.v.put_group(group, stream=self.comm, engine=engine)

    def apply(self, group, engine=None):
        """COMP + PUT of ``group``: with engine "fused" (the default where the successor reads the halos at once)
        the kernel itself pushes its edge rows into the neighbours' windows; otherwise kernel, then PUT."""
        engine = engine or {"kernel": "fused"}.get(self.engines.get(group, "kernel"), "copy")
        if engine == "fused":
            self.v.apply_group_put(group)
        else:
            self.v.apply_group(group, with_halo=False)
            self.update_async(group, engine)

    def wait(self, group):
        self.v.wait_group(group)

    def update(self, group, engine="kernel"):
        self.v.put_group(group, engine=engine)
        self.v.wait_group(group)
