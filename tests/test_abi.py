"""The C-ABI library loads without a GPU and exports every symbol include/absinthe_b200.h declares."""

import ctypes
import os
import re

import numpy as np
import pytest

from absinthe_b200 import launcher

HEADER = os.path.join(os.path.dirname(__file__), "..", "include", "absinthe_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(absinthe_\w+)\s*\(", text)))


def test_header_and_python_agree():
    assert declared_symbols() == sorted(launcher.SYMBOLS)


def test_every_symbol_is_exported(launcher_lib):
    for name in declared_symbols():
        assert hasattr(launcher_lib, name), name


def test_abi_version(launcher_lib):
    assert launcher_lib.absinthe_abi_version() == 1


def test_fill_rand_is_glibc_rand(launcher_lib):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(0)
    want = [float(libc.rand()) for _ in range(50)]
    got = np.empty(50)
    launcher_lib.absinthe_fill_rand(got.ctypes.data, 50, 0, 1)
    assert got.tolist() == want
    more = launcher.reference_inputs(2, (2, 3, 4), seed=0)
    libc.srand(0)
    flat = [float(libc.rand()) for _ in range(48)]
    assert more[0].ravel().tolist() == flat[:24] and more[1].ravel().tolist() == flat[24:]


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(launcher.LauncherError, match="no CPU fallback"):
        launcher.load_library(str(tmp_path / "libabsent.so"))


def test_load_rejects_bad_arguments(launcher_lib):
    handle = ctypes.c_void_p()
    rc = launcher_lib.absinthe_variant_load(None, None, 0, 0, ctypes.byref(handle))
    assert rc != 0 and b"null argument" in launcher_lib.absinthe_last_error()
