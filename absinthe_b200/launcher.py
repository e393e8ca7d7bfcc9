"""ctypes host of the C-ABI launcher (``include/absinthe_b200.h``).

``load_library()`` dlopens ``absinthe_b200/lib/libabsinthe_b200.so`` -- it raises
if the library has not been built; there is no CPU fallback.  ``Variant`` wraps
one loaded (descriptor, cubin) pair: bind torch CUDA tensors, apply the fused
sequence, run the reference's timing protocol.  PyTorch is only the allocator
and the stream provider here.
"""

import ctypes
import os

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
LIB_PATH = os.path.join(LIB_DIR, "libabsinthe_b200.so")

_lib = None

# every symbol include/absinthe_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "absinthe_abi_version", "absinthe_last_error", "absinthe_device_count",
    "absinthe_variant_load", "absinthe_variant_free", "absinthe_variant_info",
    "absinthe_variant_field", "absinthe_variant_field_bytes", "absinthe_variant_launches",
    "absinthe_variant_fused_halo", "absinthe_variant_bind", "absinthe_variant_apply",
    "absinthe_variant_apply_graph", "absinthe_variant_apply_group", "absinthe_variant_halo_group",
    "absinthe_variant_group_outputs",
    "absinthe_variant_group_output", "absinthe_variant_field_ptr", "absinthe_variant_run",
    "absinthe_variant_apply_host", "absinthe_make_periodic", "absinthe_fill_rand", "absinthe_flush_l2",
    "absinthe_halo_pack", "absinthe_halo_unpack", "absinthe_make_periodic_ik",
    "absinthe_variant_pack_group", "absinthe_variant_unpack_group",
    "absinthe_variant_apply_host_async", "absinthe_variant_wait_host", "absinthe_compare_fields",
    "absinthe_variant_host_interior",
    "absinthe_comm_create", "absinthe_comm_connect", "absinthe_variant_put_group", "absinthe_variant_wait_group",
    "absinthe_variant_apply_group_put",
    "absinthe_comm_destroy",
]
COMM_HANDLE_BYTES = 64


class LauncherError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen the launcher; fails loudly when the CUDA extension is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise LauncherError(
            "absinthe_b200 launcher not built: %s is missing (run `python -c 'import __graft_entry__ "
            "as g; g.build()'` or absinthe_b200.build.build_launcher()); there is no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    c = ctypes
    lib.absinthe_last_error.restype = c.c_char_p
    lib.absinthe_variant_load.argtypes = [c.c_char_p, c.c_void_p, c.c_size_t, c.c_int, c.POINTER(c.c_void_p)]
    lib.absinthe_variant_free.argtypes = [c.c_void_p]
    lib.absinthe_variant_free.restype = None
    lib.absinthe_variant_info.argtypes = [c.c_void_p, c.POINTER(c.c_int), c.POINTER(c.c_int),
                                          c.POINTER(c.c_int), c.POINTER(c.c_int)]
    lib.absinthe_variant_field.argtypes = [c.c_void_p, c.c_int, c.c_int]
    lib.absinthe_variant_field.restype = c.c_char_p
    lib.absinthe_variant_field_bytes.argtypes = [c.c_void_p]
    lib.absinthe_variant_field_bytes.restype = c.c_size_t
    lib.absinthe_variant_launches.argtypes = [c.c_void_p]
    lib.absinthe_variant_fused_halo.argtypes = [c.c_void_p]
    lib.absinthe_variant_bind.argtypes = [c.c_void_p, c.POINTER(c.c_void_p), c.POINTER(c.c_void_p)]
    lib.absinthe_variant_apply.argtypes = [c.c_void_p, c.c_void_p]
    lib.absinthe_variant_apply_graph.argtypes = [c.c_void_p, c.c_void_p]
    lib.absinthe_variant_apply_group.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_void_p]
    lib.absinthe_variant_halo_group.argtypes = [c.c_void_p, c.c_int, c.c_void_p]
    lib.absinthe_variant_group_outputs.argtypes = [c.c_void_p, c.c_int]
    lib.absinthe_variant_group_output.argtypes = [c.c_void_p, c.c_int, c.c_int]
    lib.absinthe_variant_group_output.restype = c.c_char_p
    lib.absinthe_variant_field_ptr.argtypes = [c.c_void_p, c.c_char_p]
    lib.absinthe_variant_field_ptr.restype = c.c_void_p
    lib.absinthe_variant_run.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int,
                                         c.POINTER(c.c_float), c.POINTER(c.c_float)]
    lib.absinthe_variant_apply_host.argtypes = [c.c_void_p, c.POINTER(c.c_void_p), c.POINTER(c.c_void_p),
                                                c.c_void_p]
    lib.absinthe_variant_apply_host_async.argtypes = lib.absinthe_variant_apply_host.argtypes
    lib.absinthe_variant_wait_host.argtypes = [c.c_void_p]
    lib.absinthe_variant_host_interior.argtypes = [c.c_void_p, c.c_int]
    lib.absinthe_compare_fields.argtypes = [c.c_void_p, c.c_void_p] + [c.c_int] * 6 + [
        c.c_double, c.POINTER(c.c_ulonglong), c.POINTER(c.c_ulonglong), c.c_void_p]
    lib.absinthe_make_periodic.argtypes = [c.c_void_p] + [c.c_int] * 6 + [c.c_void_p]
    lib.absinthe_make_periodic_ik.argtypes = [c.c_void_p] + [c.c_int] * 6 + [c.c_void_p]
    lib.absinthe_fill_rand.argtypes = [c.c_void_p, c.c_size_t, c.c_uint, c.c_int]
    lib.absinthe_fill_rand.restype = None
    lib.absinthe_flush_l2.argtypes = [c.c_void_p]
    lib.absinthe_halo_pack.argtypes = [c.c_void_p, c.c_void_p] + [c.c_int] * 8 + [c.c_void_p]
    lib.absinthe_halo_unpack.argtypes = [c.c_void_p, c.c_void_p] + [c.c_int] * 8 + [c.c_void_p]
    lib.absinthe_variant_pack_group.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_void_p, c.c_int, c.c_void_p]
    lib.absinthe_variant_unpack_group.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p]
    lib.absinthe_comm_create.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_char_p]
    lib.absinthe_comm_connect.argtypes = [c.c_void_p, c.c_char_p, c.c_char_p]
    lib.absinthe_variant_put_group.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_void_p]
    lib.absinthe_variant_wait_group.argtypes = [c.c_void_p, c.c_int, c.c_void_p]
    lib.absinthe_comm_destroy.argtypes = [c.c_void_p]
    lib.absinthe_variant_apply_group_put.argtypes = [c.c_void_p, c.c_int, c.c_void_p]
    if path == LIB_PATH:
        _lib = lib
    return lib


def check(lib, code, what):
    if code != 0:
        raise LauncherError("%s failed (%d): %s" % (what, code, (lib.absinthe_last_error() or b"").decode()))


def _stream_handle(stream, device=None):
    """The raw stream: ``None`` is torch's current stream of ``device`` (of the current device when not given)."""
    if stream is None:
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    if isinstance(stream, int):
        return ctypes.c_void_p(stream)
    return ctypes.c_void_p(stream.cuda_stream)


def _pointer_array(values):
    arr = (ctypes.c_void_p * max(len(values), 1))()
    for n, v in enumerate(values):
        arr[n] = v
    return arr


class Variant:
    """One generated implementation variant, loaded on one GPU."""

    def __init__(self, descriptor_path, cubin_path, device=0):
        self.lib = load_library()
        with open(descriptor_path, "rb") as handle:
            descriptor = handle.read()
        with open(cubin_path, "rb") as handle:
            cubin = handle.read()
        self._cubin = ctypes.create_string_buffer(cubin, len(cubin))
        handle_ = ctypes.c_void_p()
        check(self.lib, self.lib.absinthe_variant_load(descriptor, self._cubin, len(cubin), device,
                                                       ctypes.byref(handle_)), "absinthe_variant_load")
        self.handle = handle_
        self.device = device
        dims = (ctypes.c_int * 6)()
        n_in, n_out, n_groups = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(self.lib, self.lib.absinthe_variant_info(self.handle, dims, ctypes.byref(n_in),
                                                       ctypes.byref(n_out), ctypes.byref(n_groups)), "info")
        self.X, self.Y, self.Z, self.HX, self.HY, self.HZ = list(dims)
        self.inputs = [self.lib.absinthe_variant_field(self.handle, 0, n).decode() for n in range(n_in.value)]
        self.outputs = [self.lib.absinthe_variant_field(self.handle, 1, n).decode() for n in range(n_out.value)]
        self.n_groups = n_groups.value
        self.shape = (self.Z + 2 * self.HZ, self.Y + 2 * self.HY, self.X + 2 * self.HX)
        self.field_bytes = self.lib.absinthe_variant_field_bytes(self.handle)
        self.points = self.X * self.Y * self.Z
        self._bound = None
        self._host = []

    @classmethod
    def from_program(cls, program, template="template_tiling.cu", device=0):
        from .generator import build_variant
        desc, cubin = build_variant(program, template)
        return cls(desc, cubin, device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.absinthe_variant_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- buffers -------------------------------------------------------------------------
    def empty_field(self):
        import torch
        return torch.zeros(self.shape, dtype=torch.float64, device="cuda:%d" % self.device)

    def bind(self, inputs, outputs):
        """inputs / outputs: dict name -> CUDA float64 tensor of ``self.shape`` (or lists in order)."""
        if isinstance(inputs, dict):
            inputs = [inputs[n] for n in self.inputs]
        if isinstance(outputs, dict):
            outputs = [outputs[n] for n in self.outputs]
        for t in list(inputs) + list(outputs):
            assert t.is_cuda and t.is_contiguous() and tuple(t.shape) == self.shape, "bad field tensor"
            assert t.dtype.is_floating_point and t.element_size() == 8
        check(self.lib, self.lib.absinthe_variant_bind(
            self.handle, _pointer_array([t.data_ptr() for t in inputs]),
            _pointer_array([t.data_ptr() for t in outputs])), "absinthe_variant_bind")
        self._bound = (list(inputs), list(outputs))     # keep the tensors alive

    @property
    def launches(self):
        return self.lib.absinthe_variant_launches(self.handle)

    # -- execution -----------------------------------------------------------------------
    def apply(self, stream=None, graph=False):
        fn = self.lib.absinthe_variant_apply_graph if graph else self.lib.absinthe_variant_apply
        check(self.lib, fn(self.handle, _stream_handle(stream, self.device)), "absinthe_variant_apply")

    def apply_group(self, group, with_halo=True, stream=None):
        check(self.lib, self.lib.absinthe_variant_apply_group(self.handle, group, int(with_halo),
                                                              _stream_handle(stream, self.device)), "apply_group")

    def halo_group(self, group, stream=None):
        check(self.lib, self.lib.absinthe_variant_halo_group(self.handle, group, _stream_handle(stream, self.device)),
              "halo_group")

    def pack_group(self, group, low, high, local_faces=True, stream=None):
        check(self.lib, self.lib.absinthe_variant_pack_group(self.handle, group, low.data_ptr(), high.data_ptr(),
                                                             int(local_faces), _stream_handle(stream, self.device)), "pack_group")

    def unpack_group(self, group, low, high, stream=None):
        check(self.lib, self.lib.absinthe_variant_unpack_group(self.handle, group, low.data_ptr(), high.data_ptr(),
                                                               _stream_handle(stream, self.device)), "unpack_group")

    # -- multi-GPU exchange (P2P windows, see include/absinthe_b200.h) ------------------------
    def comm_create(self, rank, nranks):
        """Allocate this rank's exchange window; returns its 64-byte IPC handle (to be handed to the neighbours)."""
        handle = ctypes.create_string_buffer(COMM_HANDLE_BYTES)
        check(self.lib, self.lib.absinthe_comm_create(self.handle, rank, nranks, handle), "absinthe_comm_create")
        return handle.raw

    def comm_connect(self, prev_handle, next_handle):
        check(self.lib, self.lib.absinthe_comm_connect(self.handle, prev_handle, next_handle), "absinthe_comm_connect")

    def put_group(self, group, stream=None, engine="kernel"):
        """PUT of the group's edge rows: ``engine`` "kernel" (one small kernel, lowest latency) or "copy" (copy
        engines, takes nothing from a compute kernel running at the same time)."""
        check(self.lib, self.lib.absinthe_variant_put_group(self.handle, group, {"kernel": 0, "copy": 1}[engine],
                                                            _stream_handle(stream, self.device)),
              "absinthe_variant_put_group")

    def apply_group_put(self, group, stream=None):
        """The group's kernel with its j-wrapped copies stored straight into the neighbours' windows, then the flags."""
        check(self.lib, self.lib.absinthe_variant_apply_group_put(self.handle, group, _stream_handle(stream, self.device)),
              "absinthe_variant_apply_group_put")

    def wait_group(self, group, stream=None):
        check(self.lib, self.lib.absinthe_variant_wait_group(self.handle, group, _stream_handle(stream, self.device)),
              "absinthe_variant_wait_group")

    def group_outputs(self, group):
        n = self.lib.absinthe_variant_group_outputs(self.handle, group)
        return [self.lib.absinthe_variant_group_output(self.handle, group, k).decode() for k in range(n)]

    def field_ptr(self, name):
        return self.lib.absinthe_variant_field_ptr(self.handle, name.encode())

    def run(self, runs, flush=True, training=False, stream=None):
        """The reference's measurement loop; returns (total_ms[runs], halo_ms[runs])."""
        total = (ctypes.c_float * runs)()
        halo = (ctypes.c_float * runs)()
        check(self.lib, self.lib.absinthe_variant_run(self.handle, _stream_handle(stream, self.device), runs, int(flush),
                                                      int(training), total, halo), "absinthe_variant_run")
        return list(total), list(halo)

    def apply_host(self, host_inputs, host_outputs, stream=None, wait=True):
        """End to end with HOST (pinned) tensors: H2D, apply, D2H.  ``wait=False`` returns while the
        copies are in flight on the variant's private copy streams; ``wait_host`` then blocks for the
        oldest step in flight.  Two steps may be in flight (with their own host output tensors)."""
        fn = self.lib.absinthe_variant_apply_host if wait else self.lib.absinthe_variant_apply_host_async
        # ``None`` entries keep the device-resident field (constants bound once)
        check(self.lib, fn(self.handle, _pointer_array([0 if t is None else t.data_ptr() for t in host_inputs]),
                           _pointer_array([t.data_ptr() for t in host_outputs]), _stream_handle(stream, self.device)),
              "apply_host")
        self._host = (self._host + [(host_inputs, host_outputs)])[-2:]   # keep the buffers of steps in flight alive

    def host_interior(self, enable=True):
        """Upload only the interiors of host inputs; the library rebuilds their periodic halos on the device."""
        check(self.lib, self.lib.absinthe_variant_host_interior(self.handle, int(enable)), "host_interior")

    def wait_host(self):
        check(self.lib, self.lib.absinthe_variant_wait_host(self.handle), "wait_host")


# ---- field helpers -----------------------------------------------------------------------

def make_periodic(tensor, dims, stream=None, ik_only=False):
    import torch
    lib = load_library()
    fn = lib.absinthe_make_periodic_ik if ik_only else lib.absinthe_make_periodic
    with torch.cuda.device(tensor.device):       # the field helpers launch on the current device: the tensor's
        check(lib, fn(tensor.data_ptr(), *dims, _stream_handle(stream, tensor.device)), "absinthe_make_periodic")


def compare_fields(expected, actual, dims, tolerance=1e-7, stream=None):
    """(errors, matches) over the interior with the reference's relative test (ref template_tiling.cpp:383-392)."""
    lib = load_library()
    import torch
    errors, matches = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    with torch.cuda.device(expected.device):
        check(lib, lib.absinthe_compare_fields(expected.data_ptr(), actual.data_ptr(), *dims, tolerance,
                                               ctypes.byref(errors), ctypes.byref(matches),
                                               _stream_handle(stream, expected.device)), "absinthe_compare_fields")
    return errors.value, matches.value


def flush_l2(stream=None):
    lib = load_library()
    check(lib, lib.absinthe_flush_l2(_stream_handle(stream)), "absinthe_flush_l2")


def reference_inputs(count, shape, seed=0):
    """Host arrays filled like the reference's ``array()`` (srand(seed); rand() per cell, field by field)."""
    import numpy as np
    lib = load_library()
    fields = []
    for n in range(count):
        a = np.empty(shape, dtype=np.float64)
        lib.absinthe_fill_rand(a.ctypes.data, a.size, seed, 1 if n == 0 else 0)
        fields.append(a)
    return fields
