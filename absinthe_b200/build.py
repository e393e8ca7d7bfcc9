"""Build the native launcher in-tree (``absinthe_b200/lib/libabsinthe_b200.so``) for sm_100a."""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(HERE, "csrc", "launcher.cu")
HEADER = os.path.join(HERE, "..", "include", "absinthe_b200.h")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libabsinthe_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def build_launcher(force=False):
    """Compile the launcher when it is missing or older than its sources; returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    newest = max(os.path.getmtime(SOURCE), os.path.getmtime(HEADER))
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    tmp = LIB_PATH + ".%d.tmp" % os.getpid()
    subprocess.check_call(["nvcc"] + NVCC_FLAGS + [SOURCE, "-o", tmp])
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_launcher(force=True))
